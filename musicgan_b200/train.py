"""`train(run_name, input_dataset_path, output_dir)` with the reference's signature and recipe (train.py:18-278):
WGAN-GP, n_critic 5, Adam(1e-3, betas (0, 0.9)), progressive growing over 8 stages.  Hyper-parameters that the
reference hard-codes are keyword-only extras with the reference's values as defaults.

The step bodies run as CUDA graphs (graphed.GraphedSteps: critic step and generator step captured once per growth stage,
the fade-in weight alpha fed as a device scalar, re-captured at every `next_layer()`); losses and critic outputs stay on
the device in a sliding window that is read back every `log_every` iterations -- the reference synchronises six times per
iteration (train.py:180-186,218-221).  `cuda_graphs=False` runs the same steps eagerly.

Launched under torchrun the batch is sharded over the ranks (each rank draws `batch_size` samples, so the global batch
is world * batch_size) and the gradients of each optimiser step are averaged with one flat-bucket all-reduce
(parallel.FlatGradBucket).  mlflow logging is used when the package is present, silently skipped otherwise."""
from __future__ import annotations

from os import mkdir
from os.path import exists, isdir

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
          max_iterations: int = None, seed: int = None, cuda_graphs: bool = True, log_every: int = 20,
          train_lengths=(50000, 100000, 150000, 200000, 250000, 300000, 350000),
          fadein_lengths=(1, 25000, 37500, 50000, 62500, 75000, 87500, 100000)) -> None:
    assert isdir(input_dataset_path), \
        f"\"{input_dataset_path}\" doesn't exist or is not a directory"
    parallel.ensure_dir(output_dir, f"\"{output_dir}\" is not a directory !")      # every rank of a torchrun launch gets here

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
    optim_gen = th.optim.Adam(gen.parameters(), lr=gen_lr, betas=betas, fused=True, capturable=cuda_graphs)
    optim_disc = th.optim.Adam(disc.parameters(), lr=disc_lr, betas=betas, fused=True, capturable=cuda_graphs)
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

    grower = Grower(n_grow=7, fadein_lengths=list(fadein_lengths), train_lengths=list(train_lengths))      # train.py:101-109
    saver = Saver(output_dir, save_every=save_every, rand_channels=rand_channels, rand_height=height, rand_width=width)

    window = 20
    dev = th.device("cuda", th.cuda.current_device())
    # sliding 20-iteration window of [e_tp, e_tn, d_loss, grad_pen | e_gen, g_loss] (train.py:120-127) kept ON the device
    ring = th.zeros(window, 6, device=dev)
    order_d, order_g = th.tensor([2, 3, 0, 1], device=dev), th.tensor([1, 0], device=dev)      # graph stats -> window columns
    shown = [0.] * 6
    last = [0.] * 6

    def capture():
        from .graphed import GraphedSteps
        return GraphedSteps(gen, disc, optim_gen, optim_disc, batch_size, rand_channels, grower.target_size, grower.alpha,
                            bucket_d=bucket_d if ws > 1 else None, bucket_g=bucket_g if ws > 1 else None, preserve_state=True)

    graphed = capture() if cuda_graphs else None
    iter_idx, gen_idx = 0, 0
    for e in range(nb_epoch):
        if sampler is not None:
            sampler.set_epoch(e)
        bar = tqdm(DevicePrefetcher(loader, dev), total=len(loader), disable=rank != 0)
        for x_real in bar:
            # the reference normalises / resizes on the CPU and then uploads (train.py:139-140); here the fp64 chunk is
            # uploaded once (one step ahead, on a copy stream) and everything else happens on the GPU
            x_real = grower.scale_transform(x_real.to(th.float))
            alpha = grower.alpha
            do_gen = iter_idx % n_critic == 0

            if graphed is not None:
                graphed.set_alpha(alpha)
                d_stats = graphed.critic_step(x_real)            # [d_loss, grad_pen, mean D(real), mean D(fake)]
                ring[iter_idx % window, :4].copy_(d_stats[order_d])
                if do_gen:
                    g_stats = graphed.generator_step()           # [g_loss, mean D(fake)]
                    ring[gen_idx % window, 4:].copy_(g_stats[order_g])
            else:
                z = th.randn(batch_size, rand_channels, height, width, device="cuda")
                d_loss, gp, out_real, out_fake = train_step.critic_step(gen, disc, None, z, x_real, alpha, step=False)
                bucket_d.sync()
                optim_disc.step()
                ring[iter_idx % window, :4].copy_(th.stack([out_real.mean(), out_fake.mean(), d_loss, gp]))
                if do_gen:
                    z = th.randn(batch_size, rand_channels, height, width, device="cuda")
                    gen_loss, out_fake = train_step.generator_step(gen, disc, None, z, alpha, step=False)
                    bucket_g.sync()
                    optim_gen.step()
                    ring[gen_idx % window, 4:].copy_(th.stack([out_fake.mean(), gen_loss]))
            gen_idx += int(do_gen)

            if iter_idx % log_every == 0:                         # ONE device->host sync every log_every iterations
                filled = min(iter_idx + 1, window)
                both = th.cat([ring.sum(0) / filled, ring[iter_idx % window, :4], ring[(gen_idx - 1) % window, 4:]]).tolist()
                shown, last = both[:6], both[6:]
                if gen_idx < window:                              # the generator columns fill five times more slowly
                    g_mean = (ring[:, 4:].sum(0) / max(min(gen_idx, window), 1)).tolist()
                    shown[4], shown[5] = g_mean
            bar.set_description(
                f"Epoch {e:02} [{saver.curr_save:03}: {saver.save_counter:03}], "
                f"disc_l = {shown[2]:.4f}, gen_l = {shown[5]:.2f}, grad_p = {shown[3]:.4f}, "
                f"e_tp = {shown[0]:.2f}, e_tn = {shown[1]:.2f}, e_gen = {shown[4]:.2f}, alpha = {alpha:.3f}")
            if use_mlflow and iter_idx % 200 == 0:
                mlflow.log_metrics(step=gen.curr_layer, metrics={"disc_loss": last[2], "gen_loss": last[5],
                                                                 "batch_tp_error": last[0], "batch_tn_error": last[1]})
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
                if graphed is not None:                           # new modules, new resolution: capture this stage's graphs
                    graphed = None
                    th.cuda.empty_cache()
                    graphed = capture()
                bar.write(f"\nNext layer, {gen.curr_layer} / {gen.down_sample}, curr_save = {saver.curr_save}")
            if max_iterations is not None and iter_idx >= max_iterations:
                return
