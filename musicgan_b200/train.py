"""`train(run_name, input_dataset_path, output_dir)` with the reference's signature and recipe (train.py:18-278):
WGAN-GP, n_critic 5, Adam(1e-3, betas (0, 0.9)), progressive growing over 8 stages.  Hyper-parameters that the
reference hard-codes are keyword-only extras with the reference's values as defaults.

Launched under torchrun the batch is sharded over the ranks (each rank draws `batch_size` samples, so the global batch
is world * batch_size) and the gradients of each optimiser step are averaged with one flat-bucket all-reduce
(parallel.FlatGradBucket).  mlflow logging is used when the package is present, silently skipped otherwise."""
from __future__ import annotations

from os import mkdir
from os.path import exists, isdir
from statistics import mean

import torch as th
from torch.utils.data import DataLoader
from tqdm import tqdm

from . import audio, networks, parallel, train_step
from .utils import DevicePrefetcher, Grower, Saver

try:                                    # observability only, never on the hot path
    import mlflow
except ImportError:                     # pragma: no cover
    mlflow = None


def train(run_name: str, input_dataset_path: str, output_dir: str, *,
          batch_size: int = 6, nb_epoch: int = 1000, rand_channels: int = 32, disc_lr: float = 1e-3,
          gen_lr: float = 1e-3, betas=(0.0, 0.9), n_critic: int = 5, save_every: int = 1000, num_workers: int = 6,
          max_iterations: int = None, seed: int = None) -> None:
    assert isdir(input_dataset_path), \
        f"\"{input_dataset_path}\" doesn't exist or is not a directory"
    if not exists(output_dir):
        mkdir(output_dir)
    elif exists(output_dir) and not isdir(output_dir):
        raise NotADirectoryError(f"\"{output_dir}\" is not a directory !")

    rank, ws = parallel.world()
    if seed is not None:
        th.manual_seed(seed)            # identical initial weights on every rank
    height = width = 2
    gen = networks.Generator(rand_channels, end_layer=0).cuda()
    disc = networks.Discriminator(start_layer=7).cuda()
    if ws > 1:                          # same initial parameters everywhere: broadcast rank 0's
        import torch.distributed as dist
        for p in list(gen.parameters()) + list(disc.parameters()):
            dist.broadcast(p.data, src=0)
        th.manual_seed(1000 + rank + (seed or 0))     # per-rank latent / epsilon streams
    # fused=True: one multi-tensor update kernel per optimizer step instead of ~100 tiny foreach launches (same fp32 math)
    optim_gen = th.optim.Adam(gen.parameters(), lr=gen_lr, betas=betas, fused=True)
    optim_disc = th.optim.Adam(disc.parameters(), lr=disc_lr, betas=betas, fused=True)
    bucket_g, bucket_d = parallel.FlatGradBucket(gen.parameters()), parallel.FlatGradBucket(disc.parameters())

    dataset = audio.AudioDataset(input_dataset_path)
    sampler = None
    if ws > 1:
        from torch.utils.data.distributed import DistributedSampler
        sampler = DistributedSampler(dataset, num_replicas=ws, rank=rank, shuffle=True, drop_last=True)
    loader = DataLoader(dataset, batch_size=batch_size, shuffle=sampler is None, sampler=sampler,
                        num_workers=num_workers, drop_last=True, pin_memory=True)

    use_mlflow = mlflow is not None and rank == 0
    if use_mlflow:
        mlflow.set_experiment("music_gan")
        mlflow.start_run(run_name=run_name)
        mlflow.log_params({"input_dataset": input_dataset_path, "nb_sample": len(dataset), "output_dir": output_dir,
                           "rand_channels": rand_channels, "nb_epoch": nb_epoch, "batch_size": batch_size, "world_size": ws,
                           "disc_lr": disc_lr, "gen_lr": gen_lr, "betas": betas, "sample_rate": audio.SAMPLE_RATE,
                           "width": width, "height": height})

    grower = Grower(n_grow=7, fadein_lengths=[1, 25000, 37500, 50000, 62500, 75000, 87500, 100000],
                    train_lengths=[50000, 100000, 150000, 200000, 250000, 300000, 350000])
    saver = Saver(output_dir, save_every=save_every, rand_channels=rand_channels, rand_height=height, rand_width=width)

    window = 20
    hist = {k: [0.] * window for k in ("tp", "tn", "gen", "d_loss", "gp", "g_loss")}

    def push(key, value):
        del hist[key][0]
        hist[key].append(value)

    iter_idx = 0
    gen_loss = th.zeros((), device="cuda")
    for e in range(nb_epoch):
        if sampler is not None:
            sampler.set_epoch(e)
        bar = tqdm(DevicePrefetcher(loader, th.device("cuda", th.cuda.current_device())), total=len(loader), disable=rank != 0)
        for x_real in bar:
            # the reference normalises / resizes on the CPU and then uploads (train.py:139-140); here the fp64 chunk is
            # uploaded once (one step ahead, on a copy stream) and everything else happens on the GPU
            x_real = grower.scale_transform(x_real.to(th.float))
            alpha = grower.alpha

            z = th.randn(batch_size, rand_channels, height, width, device="cuda")
            d_loss, gp, out_real, out_fake = train_step.critic_step(gen, disc, None, z, x_real, alpha, step=False)
            bucket_d.sync()
            optim_disc.step()
            stats = th.stack([out_real.mean(), out_fake.mean(), d_loss, gp]).tolist()     # ONE device->host sync
            push("tp", stats[0]); push("tn", stats[1]); push("d_loss", stats[2]); push("gp", stats[3])

            if iter_idx % n_critic == 0:
                z = th.randn(batch_size, rand_channels, height, width, device="cuda")
                gen_loss, out_fake = train_step.generator_step(gen, disc, None, z, alpha, step=False)
                bucket_g.sync()
                optim_gen.step()
                g_stats = th.stack([out_fake.mean(), gen_loss]).tolist()
                push("gen", g_stats[0]); push("g_loss", g_stats[1])

            bar.set_description(
                f"Epoch {e:02} [{saver.curr_save:03}: {saver.save_counter:03}], "
                f"disc_l = {mean(hist['d_loss']):.4f}, gen_l = {mean(hist['g_loss']):.2f}, grad_p = {mean(hist['gp']):.4f}, "
                f"e_tp = {mean(hist['tp']):.2f}, e_tn = {mean(hist['tn']):.2f}, e_gen = {mean(hist['gen']):.2f}, alpha = {alpha:.3f}")
            if use_mlflow and iter_idx % 200 == 0:
                mlflow.log_metrics(step=gen.curr_layer, metrics={"disc_loss": stats[2], "gen_loss": hist["g_loss"][-1],
                                                                 "batch_tp_error": stats[0], "batch_tn_error": stats[1]})
            if rank == 0:
                saver.request_save(gen, disc, optim_gen, optim_disc, alpha)
            iter_idx += 1

            # ProGAN: next layer.  The schedule counts the samples seen by ONE replica, like the reference (so a run on N
            # GPUs sees N times more data per stage unless the caller scales the lengths)
            if grower.grow(batch_size) and gen.growing:
                gen.next_layer()
                disc.next_layer()
                optim_gen.add_param_group({"params": gen.end_block_params(), "lr": gen_lr, "betas": betas})
                optim_disc.add_param_group({"params": disc.start_block_parameters(), "lr": disc_lr, "betas": betas})
                if ws > 1:
                    import torch.distributed as dist
                    for p in list(gen.end_block_params()) + list(disc.start_block_parameters()):
                        dist.broadcast(p.data, src=0)
                bucket_g.rebuild(gen.parameters())
                bucket_d.rebuild(disc.parameters())
                bar.write(f"\nNext layer, {gen.curr_layer} / {gen.down_sample}, curr_save = {saver.curr_save}")
            if max_iterations is not None and iter_idx >= max_iterations:
                return
