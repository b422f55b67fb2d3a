"""CPU: host-side schedule logic of utils.Saver / utils.Grower against the reference classes themselves (imported from
/root/reference when present -- the authoring container; skipped on a box without it) and against pinned sequences."""
import os
import sys
import types

import pytest
import torch


def _ref_utils():
    if not os.path.isdir("/root/reference/music_gan"):
        pytest.skip("reference tree not present")
    for n in ("mlflow", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    sys.dont_write_bytecode = True
    from music_gan import utils as ref_utils
    return ref_utils


class _Dummy:
    def state_dict(self):
        return {"w": torch.zeros(1)}


def _schedule(saver, calls):
    out = []
    for _ in range(calls):
        fired = saver.request_save(_Dummy(), _Dummy(), _Dummy(), _Dummy(), 1.0)
        out.append((bool(fired), saver.curr_save, saver.save_counter))
    return out


# reference utils.py:209-242: count first, save when count % save_every == 0; curr_save = last saved index; counter mod
PINNED_EVERY3 = [(False, -1, 1), (False, -1, 2), (True, 0, 0), (False, 0, 1), (False, 0, 2), (True, 1, 0), (False, 1, 1)]


def test_saver_schedule_pinned(tmp_path):
    from musicgan_b200.utils import Saver
    got = _schedule(Saver(str(tmp_path), save_every=3, rand_channels=32), 7)
    assert got == PINNED_EVERY3
    assert sorted(os.listdir(tmp_path)) == sorted(f"{k}_{i}.pt" for i in (0, 1) for k in ("disc", "gen", "optim_disc", "optim_gen"))


@pytest.mark.parametrize("every", [1, 3, 5])
def test_saver_schedule_matches_reference_class(tmp_path, every):
    ref_utils = _ref_utils()
    from musicgan_b200.utils import Saver
    ref = ref_utils.Saver(str(tmp_path / "ref"), save_every=every, rand_channels=32, rand_height=2, rand_width=2) \
        if "rand_height" in ref_utils.Saver.__init__.__code__.co_varnames else None
    if ref is None:
        pytest.skip("unexpected reference Saver signature")
    # the reference writes models and matplotlib previews when it fires: replace both private hooks, keep the schedule
    ref._Saver__save_models = lambda *a, **k: None
    ref._Saver__save_outputs = lambda *a, **k: None
    (tmp_path / "ours").mkdir()
    ours = Saver(str(tmp_path / "ours"), save_every=every, rand_channels=32)
    assert _schedule(ours, 4 * every + 2) == _schedule(ref, 4 * every + 2)


def test_grower_matches_reference_class():
    ref_utils = _ref_utils()
    from musicgan_b200.utils import Grower
    fade, lens = [1, 16, 16, 50, 60, 70, 80, 100], [8, 12, 40, 50, 60, 70, 80]
    a, b = Grower(7, fade, lens), ref_utils.Grower(7, fade, lens)
    for _ in range(120):
        assert a.alpha == b.alpha
        assert a.grow(6) == b.grow(6)


def _scale_case():
    g = torch.Generator().manual_seed(4242)
    return torch.rand(3, 2, 512, 512, generator=g, dtype=torch.float64).float() * 3.0 - 1.0


def _check_scale_transform(device, golden_dir):
    """utils.Grower.scale_transform (min-max, range, antialiased bilinear resize as two small GEMMs) against the
    REFERENCE's own Grower / torchvision Compose at every growth stage (fixture: oracle/gen_golden.py scale)."""
    import os
    import numpy as np
    from musicgan_b200.utils import Grower
    gold = np.load(os.path.join(golden_dir, "scale_transform.npz"))
    x = _scale_case().to(device)
    grower = Grower(7, [1, 2, 2, 2, 2, 2, 2, 2], [1, 2, 3, 4, 5, 6, 7])
    for stage in range(8):
        y = grower.scale_transform(x)
        size = 4 * 2 ** stage
        assert tuple(y.shape) == (3, 2, size, size) and y.dtype == torch.float32
        got = y.contiguous().view(-1)[::int(gold[f"stage{stage}_stride"])].cpu().numpy()
        np.testing.assert_allclose(got, gold[f"stage{stage}"], rtol=2e-5, atol=4e-6)
        d = gold[f"stage{stage}_digest"]
        assert abs(y.double().sum().item() - d[0]) <= 2e-5 * d[1] + 1e-6
        if stage < 7:
            while not grower.grow(1):
                pass


def test_scale_transform_matches_reference_grower_cpu(golden_dir):
    _check_scale_transform("cpu", golden_dir)


@pytest.mark.gpu
def test_scale_transform_matches_reference_grower_gpu(golden_dir):
    torch.backends.cuda.matmul.allow_tf32 = False
    _check_scale_transform("cuda", golden_dir)
