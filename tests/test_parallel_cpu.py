"""CPU: the multi-GPU host logic -- shard plans for the communication-free paths and the flat-bucket gradient
all-reduce, exercised with world_size 2 on the gloo backend (one process per rank, like torchrun on the GPU box)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from musicgan_b200 import parallel
from musicgan_b200.utils import Grower
from oracle import audio_oracle as ao


def test_shard_bounds_cover_everything_once():
    for n in [0, 1, 7, 8, 1024, 1025]:
        for ws in [1, 2, 3, 8]:
            spans = [parallel.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_dataset_shard_plan_reproduces_sequential_idx(golden_dir):
    rows = np.load(os.path.join(golden_dir, "index_plan.npz"))["rows"]       # from the reference itself
    counts = [int(r[0]) for r in rows]
    for r in rows:
        assert parallel.chunks_written(int(r[0])) == int(r[3])
    seq = ao.dataset_index_plan(counts)
    for ws in [1, 2, 3, 4]:
        for rank in range(ws):
            b, e, first = parallel.dataset_shard_plan(counts, rank, ws)
            if b < e:
                assert first == seq[b][0] == int(rows[b][2])


def test_grower_schedule_matches_reference_iterations():
    # SURVEY B.5 (measured on the reference): growth happens at these iterations for batch 6
    g = Grower(7, [1, 25000, 37500, 50000, 62500, 75000, 87500, 100000], [50000, 100000, 150000, 200000, 250000, 300000, 350000])
    assert g.alpha == 1.0                       # fadein[0] == 1
    grown = []
    for i in range(1, 233400):
        if g.grow(6):
            grown.append(i)
            assert g.alpha == pytest.approx(1.0 / [25000, 37500, 50000, 62500, 75000, 87500, 100000][len(grown) - 1])
    assert grown == [8334, 25001, 50001, 83334, 125001, 175001, 233334]
    assert g.target_size == 512


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3), torch.nn.Linear(3, 2))
        # every rank: a different shard of the same global batch; layer 2 inactive (grad None) on all ranks
        full = torch.arange(8 * 5, dtype=torch.float32).reshape(8, 5) / 10.0
        b, e = parallel.shard_bounds(8, rank, world)
        out = lin[1](lin[0](full[b:e]))
        out.pow(2).mean().backward()
        bucket = parallel.FlatGradBucket(lin.parameters())
        n = bucket.sync()
        assert n == sum(p.numel() for p in lin[:2].parameters())
        assert all(p.grad is None for p in lin[2].parameters())
        # single-process reference on the concatenated batch
        torch.manual_seed(0)
        ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3), torch.nn.Linear(3, 2))
        ref[1](ref[0](full)).pow(2).mean().backward()
        for p, q in zip(lin[:2].parameters(), ref[:2].parameters()):
            torch.testing.assert_close(p.grad, q.grad, rtol=1e-5, atol=1e-6)
        # shard plans agree across ranks (no communication needed, but check consistency with a gather)
        counts = [100_000, 131_072, 300_000, 262_151, 130_816]
        mine = torch.tensor(list(parallel.dataset_shard_plan(counts, rank, world)))
        got = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(got, mine)
        assert got[0][1] == got[1][0]
        assert int(got[1][2]) == sum(parallel.chunks_written(c) for c in counts[: int(got[1][0])])
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: 1, 1: 1}


@pytest.mark.parametrize("size", [4, 8, 16, 32, 64, 128, 256, 512])
def test_real_batch_transform_matches_torch_antialias_resize(size):
    """Grower.scale_transform == ChannelMinMaxNorm -> ChangeRange(-1,1) -> torchvision Resize (utils.py:70-82), the
    latter being F.interpolate(bilinear, antialias=True) on tensors."""
    import torch.nn.functional as F
    from musicgan_b200 import audio
    g = Grower(7, [1] * 8, [1] * 7)
    while g.target_size < size:
        g.grow(2)
    assert g.target_size == size
    x = torch.rand(2, 2, 512, 512, generator=torch.Generator().manual_seed(size))
    ref = audio.ChangeRange(-1., 1.)(audio.ChannelMinMaxNorm()(x))
    if size != 512:
        ref = F.interpolate(ref, size=(size, size), mode="bilinear", antialias=True, align_corners=False)
    got = g.scale_transform(x)
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=2e-6)


def _mk(path, q):
    from musicgan_b200 import parallel
    try:
        parallel.ensure_dir(path, "not a directory")
        q.put("ok")
    except Exception as e:      # noqa: BLE001
        q.put(type(e).__name__)


def test_ensure_dir_is_safe_between_ranks(tmp_path):
    """Every rank of a torchrun launch creates the output directory of create_dataset / train / generate: the reference's
    exists()-then-mkdir() raises FileExistsError in the loser (seen on two GPUs); ensure_dir must not, and must keep the
    reference's errors (a file in the way, a missing parent)."""
    import multiprocessing as mp
    from musicgan_b200 import parallel
    ctx = mp.get_context("spawn")
    for rep in range(3):
        target = str(tmp_path / f"out{rep}")
        q = ctx.Queue()
        procs = [ctx.Process(target=_mk, args=(target, q)) for _ in range(6)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        assert sorted(q.get() for _ in procs) == ["ok"] * 6
        assert os.path.isdir(target)
    f = tmp_path / "a_file"
    f.write_text("x")
    with pytest.raises(NotADirectoryError):
        parallel.ensure_dir(str(f), "not a directory")
    with pytest.raises(FileNotFoundError):
        parallel.ensure_dir(str(tmp_path / "missing" / "child"), "not a directory")
