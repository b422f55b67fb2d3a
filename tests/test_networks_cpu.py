"""CPU: drop-in module surface -- state-dict keys / shapes / registration order equal the reference's
(oracle.networks_oracle.state_shapes, itself asserted against the reference modules by gen_golden_networks),
growth bookkeeping, the loud failure on CPU tensors, and the oracle against the reference goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import networks_oracle as no
from oracle.gen_golden_networks import CASES, SMALL_CASES, case_inputs


@pytest.mark.parametrize("stage", [0, 1, 4, 7])
def test_state_dict_layout_matches_reference(stage):
    from musicgan_b200 import networks
    gen, disc = networks.Generator(32, 0), networks.Discriminator(7)
    for _ in range(stage):
        gen.next_layer(); disc.next_layer()
    assert [(k, tuple(v.shape)) for k, v in gen.state_dict().items()] == no.state_shapes("gen", stage)
    assert [(k, tuple(v.shape)) for k, v in disc.state_dict().items()] == no.state_shapes("disc", stage)
    assert gen.curr_layer == stage and disc.curr_layer == 7 - stage and gen.down_sample == 7
    assert len(list(gen.end_block_params())) == 2 and len(list(disc.start_block_parameters())) == 2


def test_generator_end_layer_7_has_last_end_block():
    from musicgan_b200 import networks
    gen = networks.Generator(32, end_layer=7)          # generate.py:29-32
    assert "_Generator__last_end_block.0.0.weight" in gen.state_dict()
    assert "_Discriminator__last_start_block.1.0.weight" not in networks.Discriminator(3).state_dict()


def test_cpu_forward_fails_loudly():
    from musicgan_b200 import networks
    with pytest.raises(RuntimeError, match="CUDA"):
        networks.Generator(32)(torch.randn(1, 32, 2, 2), 1.0)
    with pytest.raises(NotImplementedError):
        networks.Generator(8)


@pytest.mark.parametrize("name", SMALL_CASES)
def test_oracle_matches_reference_golden(golden_dir, name):
    gold = np.load(os.path.join(golden_dir, f"networks_{name}.npz"))
    stage, batch, alpha, z, z2, x_real, eps = case_inputs(name)
    sd_g, sd_d = no.make_state("gen", stage, 11), no.make_state("disc", stage, 12)
    d = no.d_step(sd_g, sd_d, z, x_real, eps, alpha, stage)
    np.testing.assert_allclose(d["x_fake"].contiguous().view(-1)[::int(gold["x_fake_stride"])].numpy(), gold["x_fake"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(d["out_real"].numpy(), gold["out_real"], rtol=1e-4, atol=1e-6)
    assert abs(d["gp"].item() - float(gold["gp"])) <= 1e-5 * abs(float(gold["gp"]))
    for k, v in d["grads"].items():
        if v is None:
            assert k in gold["none_d"].tolist()
            continue
        ref_norm = float(gold["dgrad_digest/" + k][2])
        assert abs(v.double().norm().item() - ref_norm) <= 1e-4 * max(ref_norm, 1e-12), k
    g = no.g_step(sd_g, sd_d, z2, alpha, stage)
    assert abs(g["loss"].item() - float(gold["g_loss"])) <= 1e-5 * max(abs(float(gold["g_loss"])), 1e-6)
