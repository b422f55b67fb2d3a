"""GPU (needs >= 2 devices; skipped otherwise): data-parallel critic step -- the gradients after the flat-bucket NCCL
all-reduce on 2 ranks, each with half of the batch, equal the fp32 CPU ORACLE's gradients on the concatenated batch
(oracle.networks_oracle: reference train.py:143-174 / discriminator.py:157-184; SURVEY 8e: no batch-coupled layer in D or
G, losses are batch means), term by term at rel-L2 <= 1e-2 like the single-GPU parity tests."""
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

        terms = {
            "real": (lambda sl: disc(x_real[sl], alpha).mean(), lambda d: no.disc_forward(d, x_real.cpu(), alpha, stage).mean()),
            "gp": (lambda sl: disc.gradient_penalty(x_real[sl], x_fake[sl], alpha, eps=eps[sl]),
                   lambda d: no.gradient_penalty(d, x_real.cpu(), x_fake.cpu(), alpha, stage, eps.cpu())),
            "critic": (lambda sl: -(disc(x_real[sl], alpha).mean() - disc(x_fake[sl], alpha).mean())
                       + disc.gradient_penalty(x_real[sl], x_fake[sl], alpha, eps=eps[sl]), None),
        }
        b, e = parallel.shard_bounds(batch, rank, world)
        bucket = parallel.FlatGradBucket(disc.parameters())
        out = {}
        for key, (ours, oracle) in terms.items():
            disc.zero_grad()
            ours(slice(b, e)).backward()             # this rank's shard only
            n = bucket.sync()                        # average over the two ranks
            assert n > 0
            mine = {k: p.grad.clone() for k, p in disc.named_parameters() if p.grad is not None}
            if oracle is not None:                   # the fp32 oracle on the WHOLE batch
                d = no._leaf(sd_d)
                oracle(d).backward()
                ref = {k: v.grad for k, v in d.items() if v.grad is not None}
            else:                                    # total: single process on the whole batch (cancelling terms: see test_networks_gpu)
                disc.zero_grad()
                ours(slice(0, batch)).backward()
                ref = {k: p.grad.cpu() for k, p in disc.named_parameters() if p.grad is not None}
            num = den = 0.0
            for k, r in ref.items():
                if k not in mine:
                    assert float(r.abs().max()) == 0.0, k
                    continue
                num += (mine[k].double().cpu() - r.double()).pow(2).sum().item()
                den += r.double().pow(2).sum().item()
            out[key] = (num / max(den, 1e-300)) ** 0.5
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduced_gradients_equal_single_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rank, args=(2, 29600 + os.getpid() % 1000, ret), nprocs=2, join=True)
    print("rel-L2 of 2-rank averaged grads: real / gp terms vs the fp32 oracle on the concatenated batch, critic total vs one process:", dict(ret))
    for v in ret.values():
        assert v["real"] <= 1e-2 and v["gp"] <= 1e-2, v
        assert v["critic"] <= 1e-2, v


def _gen_rank(rank, world, port, ckpt, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import musicgan_b200 as mg
        torch.manual_seed(7)                 # the latent of ALL clips is drawn by every rank from the same stream
        mg.generate(out_dir, 32, ckpt, nb_vec=1, nb_music=5)
    finally:
        dist.destroy_process_group()


def test_generate_shards_clips_over_two_ranks(tmp_path):
    """generate shards CLIPS with no communication -- two ranks write exactly the files (names and samples)
    that one process writes for the same latent draw (generate.py:47-65)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import musicgan_b200 as mg
    from musicgan_b200 import networks
    from musicgan_b200.audio import wavio
    torch.manual_seed(0)
    ckpt = str(tmp_path / "gen.pt")
    torch.save(networks.Generator(32, end_layer=7).state_dict(), ckpt)
    one, two = tmp_path / "one", tmp_path / "two"
    torch.manual_seed(7)
    mg.generate(str(one), 32, ckpt, nb_vec=1, nb_music=5)
    mp.spawn(_gen_rank, args=(2, 29700 + os.getpid() % 1000, ckpt, str(two)), nprocs=2, join=True)
    assert sorted(os.listdir(one)) == sorted(os.listdir(two)) == [f"sound_{i}.wav" for i in range(5)]
    for i in range(5):
        a, _ = wavio.load(str(one / f"sound_{i}.wav"))
        b, _ = wavio.load(str(two / f"sound_{i}.wav"))
        assert torch.equal(a, b), i
