"""GPU (needs >= 2 devices; skipped otherwise): data-parallel critic step -- the gradients after the flat-bucket NCCL
all-reduce on 2 ranks, each with half of the batch, equal the single-process gradients on the whole batch
(no batch-coupled layer in D or G, losses are batch means)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _rank(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from musicgan_b200 import networks, parallel
        from oracle import networks_oracle as no
        stage, alpha, batch = 3, 0.5, 4
        sd_d = no.make_state("disc", stage, 31)
        disc = networks.Discriminator(7)
        for _ in range(stage):
            disc.next_layer()
        disc.load_state_dict(sd_d)
        disc.cuda()
        g = torch.Generator().manual_seed(123)
        r = 4 * 2 ** stage
        x_real = (torch.rand(batch, 2, r, r, generator=g) * 2 - 1).cuda()
        x_fake = (torch.rand(batch, 2, r, r, generator=g) * 2 - 1).cuda()
        eps = torch.rand(batch, 1, 1, 1, generator=g).cuda()

        def grads(sl):
            disc.zero_grad()
            loss = -(disc(x_real[sl], alpha).mean() - disc(x_fake[sl], alpha).mean())
            gp = disc.gradient_penalty(x_real[sl], x_fake[sl], alpha, eps=eps[sl])
            (loss + gp).backward()

        b, e = parallel.shard_bounds(batch, rank, world)
        grads(slice(b, e))
        n = parallel.FlatGradBucket(disc.parameters()).sync()
        assert n > 0
        mine = {k: p.grad.clone() for k, p in disc.named_parameters() if p.grad is not None}
        grads(slice(0, batch))              # single-process reference on the concatenated batch (same kernels)
        num = den = 0.0
        for k, p in disc.named_parameters():
            if p.grad is None:
                assert k not in mine
                continue
            num += (mine[k].double() - p.grad.double()).pow(2).sum().item()
            den += p.grad.double().pow(2).sum().item()
        ret[rank] = (num / max(den, 1e-300)) ** 0.5
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduced_gradients_equal_single_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rank, args=(2, 29600 + os.getpid() % 1000, ret), nprocs=2, join=True)
    print("rel-L2 of 2-rank averaged grads vs single process:", dict(ret))
    assert all(v <= 2e-3 for v in ret.values())       # same bf16 kernels; only the per-sample split of the batch means differs
