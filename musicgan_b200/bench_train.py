"""bench.py's GAN-training workload (BASELINE config 2): one WGAN-GP iteration of the reference schedule --
a critic step every iteration and a generator step every 5th (train.py:189) -- at the final 512 x 512 stage,
batch 8 per GPU, random-init weights, Adam(lr 1e-3, betas (0, 0.9)) updates included.

Multi-GPU: batch-sharded replicas, gradients averaged with one flat-bucket NCCL all-reduce per optimiser step.
"""
from __future__ import annotations

import os
import time

import torch as th

# algorithmic GFLOP per sample of the reference's critic step (3 F_G + 12 F_D) and generator step (3 F_G + 3 F_D) at each
# stage of the progressive schedule (FlopCounterMode on the reference)
GFLOP_D_STEP = [0.111, 0.430, 1.391, 4.156, 11.601, 29.76, 67.233, 121.78]
GFLOP_G_STEP = [0.030, 0.158, 0.543, 1.649, 4.627, 11.89, 26.884, 48.71]


def flop_per_iter(stage: int) -> float:
    """one reference iteration = critic step + 1/5 generator step (train.py:189)"""
    return (GFLOP_D_STEP[stage] + GFLOP_G_STEP[stage] / 5.0) * 1e9


def _build(stage: int, seed: int, device):
    from . import networks
    th.manual_seed(seed)
    gen, disc = networks.Generator(32, 0), networks.Discriminator(7)
    for _ in range(stage):
        gen.next_layer(); disc.next_layer()
    return gen.to(device), disc.to(device)


def run(args, rank, world, local, timed_region, ClockSampler, peaks):
    from . import _lib, train_step
    from .networks import ops
    dev = th.device("cuda", local)
    stage, batch, alpha = args.stage, args.batch, 0.5
    gen, disc = _build(stage, 0, dev)
    use_graphs = bool(int(os.environ.get("MG_GRAPHS", "1")))
    opt_g = th.optim.Adam(gen.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=use_graphs, fused=True)
    opt_d = th.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=use_graphs, fused=True)
    th.cuda.manual_seed(1000 + rank)
    g = None
    res = 4 * 2 ** stage
    n_real = 4
    reals_host = [(th.rand(batch, 2, res, res) * 2 - 1).pin_memory() for _ in range(n_real)]
    reals_dev = [x.to(dev) for x in reals_host]
    it = [0]

    from . import parallel
    bucket_d = parallel.FlatGradBucket(disc.parameters()) if world > 1 else None      # the object `train` uses
    bucket_g = parallel.FlatGradBucket(gen.parameters()) if world > 1 else None
    graphed = None
    if use_graphs:
        from .graphed import GraphedSteps
        graphed = GraphedSteps(gen, disc, opt_g, opt_d, batch, 32, res, alpha, bucket_d=bucket_d, bucket_g=bucket_g)

    def graphed_iter(x_real):
        stats = graphed.critic_step(x_real)
        loss = stats[0] + stats[1]
        if it[0] % 5 == 0:
            graphed.generator_step()
        it[0] += 1
        return loss

    def one_iter(x_real):
        nonlocal graphed
        if graphed is not None:
            return graphed_iter(x_real)
        z = th.randn(batch, 32, 2, 2, device=dev, generator=g)
        with th.no_grad():
            x_fake = gen(z, alpha)
        out_both = disc(th.cat([x_real, x_fake], dim=0), alpha)
        d_loss = -(out_both[:batch].mean() - out_both[batch:].mean())
        gp = disc.gradient_penalty(x_real, x_fake, alpha)
        gen.zero_grad(); disc.zero_grad()
        (d_loss + gp).backward()
        if world > 1:
            bucket_d.sync()
        opt_d.step()
        loss = d_loss.detach() + gp.detach()
        if it[0] % 5 == 0:
            z = th.randn(batch, 32, 2, 2, device=dev, generator=g)
            g_loss = -disc(gen(z, alpha), alpha).mean()
            gen.zero_grad(); disc.zero_grad()
            g_loss.backward()
            if world > 1:
                bucket_g.sync()
            opt_g.step()
        it[0] += 1
        return loss

    def step():
        one_iter(reals_dev[it[0] % n_real])

    with ClockSampler(local) as cs:
        ms = timed_region(step, args.steps, args.warmup, world)
    # per-kernel durations and FLOP bookkeeping: graph replays cannot carry event pairs, so the SAME iteration schedule is
    # run eagerly (5 iterations = 5 critic steps + 1 generator step) with the library's event profiler switched on
    saved, graphed = graphed, None
    prof_iters = 5
    it[0] = 0
    step(); th.cuda.synchronize()
    it[0] = 0
    _lib.profile_enable(True)
    ops.FLOPS.update(count=0.0, bytes=0.0, launches=[])
    for _ in range(prof_iters):
        # an eager iteration is host bound (~1000 launches): without a head start for the host every event pair would
        # also time the gap in which the GPU waits for the next launch to be submitted.  A spin kernel keeps the GPU
        # busy while the host queues the whole iteration behind it.
        th.cuda._sleep(int(0.06 * 1.9e9))
        step()
    th.cuda.synchronize()
    prof = _lib.profile_collect(64)
    _lib.profile_enable(False)
    graphed = saved
    scale = args.steps / prof_iters
    prof = {k: (v[0] * scale, int(round(v[1] * scale))) for k, v in prof.items()}
    conv_flops_timed = ops.FLOPS["count"] * scale
    conv_bytes_timed = ops.FLOPS["bytes"] * scale
    conv_launch_list, ops.FLOPS["launches"] = ops.FLOPS["launches"], None
    value = world * args.steps / (ms * 1e-3)

    # end to end through the public API: every step's real batch comes from pinned host memory (utils.DevicePrefetcher,
    # the loader wrapper `train` uses: the upload of batch t+1 is issued on a copy stream while step t computes) and the
    # loss is read back
    import itertools
    from .utils import DevicePrefetcher
    host_loss = th.empty((), dtype=th.float32).pin_memory()
    feed = DevicePrefetcher((reals_host[i % n_real] for i in itertools.count()), dev, depth=2)

    def e2e_step():
        loss = one_iter(next(feed))
        host_loss.copy_(loss, non_blocking=True)

    e2e_steps = max(5, min(args.steps, 10))
    ms2 = timed_region(e2e_step, e2e_steps, 1, world)
    e2e_value = world * e2e_steps / (ms2 * 1e-3)

    if rank != 0:
        return None
    pk = peaks()
    conv_names = ("k_conv3x3_fprop", "k_conv3x3_dgrad", "k_conv3x3_wgrad", "k_conv3x3_split_fprop", "k_conv3x3_split_dgrad")
    conv_ms = sum(prof.get(n, (0.0, 0))[0] for n in conv_names)
    conv_launches = sum(prof.get(n, (0.0, 0))[1] for n in conv_names)
    conv_s = conv_ms * 1e-3
    tf_achieved = conv_flops_timed / conv_s / 1e12 if conv_ms > 0 else 0.0
    gb_achieved = conv_bytes_timed / conv_s / 1e9 if conv_ms > 0 else 0.0
    # which roof binds: the stage-7 layers have 16..160 channels, so most launches sit left of the ridge point
    # (2*9*Cin*Cout FLOP per 2*(Cin+Cout) bytes).  The per-launch bound is max(flops / TF, bytes / HBM); its sum over the
    # launches of a step is the time a perfect kernel would need
    t_tensor = conv_flops_timed / (pk["tf_sus"] * 1e12)
    t_hbm = conv_bytes_timed / (pk["hbm"] * 1e9)
    t_bound = sum(max(f / (pk["tf_sus"] * 1e12), b / (pk["hbm"] * 1e9)) for f, b in conv_launch_list) * scale
    hbm_bound = t_hbm >= t_tensor
    step_tf = batch * flop_per_iter(stage) * args.steps / (ms * 1e-3) / 1e12
    cpu = cpu_baseline(batch, stage=stage, budget_s=args.cpu_seconds * 2)
    torch_arm = torch_b200_numbers(stage, batch) if world == 1 and not os.environ.get("MG_BENCH_NO_TORCH") else None
    return {
        "metric": "GAN train steps/s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"ProGAN WGAN-GP iteration at {res}x{res} (BASELINE config {2 if batch == 8 and stage == 7 else 4}): critic step "
                               "every iteration + generator step every 5th, Adam updates included, alpha 0.5 (both fade paths)",
                   "stage": stage, "batch_per_gpu": batch,
                   "global_batch": batch * world, "l2": "activations of one step (>1 GB) exceed L2; 4 rotating real batches",
                   "parallelism": f"dp{world}" + (" flat-bucket NCCL all-reduce" if world > 1 else ""),
                   "cuda_graphs": use_graphs},
        "clocks": cs.summary(),
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": int(reals_host[0].numel() * 4), "d2h_bytes_per_step": 4},
        "gpu_launches": int(sum(c for _, c in prof.values())),
        "roofline": {"bound": "hbm" if hbm_bound else "tensor", "kernel": "k_conv3x3 (fprop+dgrad+wgrad, all layers)",
                     "achieved": gb_achieved if hbm_bound else tf_achieved, "peak": pk["hbm"] if hbm_bound else pk["tf_sus"],
                     "unit": "GB/s" if hbm_bound else "TFLOP/s",
                     "frac": (gb_achieved / pk["hbm"]) if hbm_bound else (tf_achieved / pk["tf_sus"]),
                     "peak_source": pk["src"] + " (sustained)",
                     # the family's dominant launch, one ncu --set full capture (profiles/r02_prof_conv_fprop_d0c1_raw.csv):
                     # D0.conv1 fprop at batch 8 moves 67.2 + 83.0 MB of DRAM traffic against 67.1 + 134.2 MB algorithmic
                     # (part of the output still sits in the 126 MB L2 when the launch ends): no wasted re-reads
                     "traffic": 150.2e6 * batch / 8.0, "traffic_source": "ncu dram__bytes read + write of D0.conv1 fprop per launch "
                                                                          "(algorithmic 201.3 MB at batch 8)",
                     "tensor": {"achieved": tf_achieved, "peak": pk["tf_sus"], "unit": "TFLOP/s", "frac": tf_achieved / pk["tf_sus"]},
                     "hbm": {"achieved": gb_achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": gb_achieved / pk["hbm"],
                             "bytes": "x read once + y written once per launch (wgrad: dy + x read), weights excluded"},
                     "per_launch_bound_frac": t_bound / conv_s if conv_ms > 0 else 0.0,
                     "kernel_timing": "event pairs around every library kernel in an eager pass of the same 5-iteration schedule "
                                      "right after the timed region, queued behind a spin kernel so the pairs do not time host "
                                      "launch gaps (graph replays cannot carry event pairs)" if use_graphs else "timed region",
                     "conv_ms_per_step": conv_ms / args.steps, "conv_launches_per_step": conv_launches / args.steps,
                     "step_algorithmic_tflops": step_tf, "step_frac": step_tf / pk["tf_sus"],
                     "kernel_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}},
        "cpu_baseline": cpu,
        "torch_b200": torch_arm,
        "samples_per_s": value * batch,
    }


def torch_b200_numbers(stage: int, batch: int, steps: int = 5, warmup: int = 3):
    """The reference's modules through stock torch on this GPU (cuDNN): the bar on the same box."""
    import gc
    from baseline import torch_b200 as tb
    th.cuda.empty_cache()
    out = {"what": "stock torch (cuDNN / ATen) restatement of the reference modules and step body, same iteration schedule, "
                   "inputs resident, CUDA events", "unit": "steps/s", "batch": batch, "stage": stage}
    for mode in ("fp32", "bf16_autocast_channels_last"):
        try:
            v, ms = tb.time_train(stage, batch, mode, steps, warmup)
            out[mode] = {"value": v, "ms_per_step": ms}
        except Exception as e:      # noqa: BLE001  (e.g. out of memory at batch 64: reported, not fatal)
            out[mode] = {"error": f"{type(e).__name__}: {str(e)[:120]}"}
        gc.collect(); th.cuda.empty_cache()
    return out


def cpu_baseline(batch: int, stage: int = 7, steps: int = 1, warmup: int = 0, budget_s: float = 30.0):
    """Oracle port of the reference step body (fp32, all host threads) at the reference's full cost (fake batch attached in
    the critic step, critic gradients computed in the generator step): `steps` iterations of the reference schedule --
    a critic step each, a generator step every 5th -- after `warmup` full-size iterations.
    value = iterations per second with the generator step weighted 1/5."""
    from oracle import networks_oracle as no
    th.set_num_threads(os.cpu_count() or 1)
    alpha, res = 0.5, 4 * 2 ** stage
    sd_g, sd_d = no.make_state("gen", stage, 1), no.make_state("disc", stage, 2)
    g = th.Generator().manual_seed(0)
    z = th.randn(batch, 32, 2, 2, generator=g)
    x_real = th.rand(batch, 2, res, res, generator=g) * 2 - 1
    eps = th.rand(batch, 1, 1, 1, generator=g)
    no.d_step(sd_g, sd_d, z[:1], x_real[:1], eps[:1], alpha, stage, attached=True)      # thread pools, allocator
    for _ in range(warmup):
        no.d_step(sd_g, sd_d, z, x_real, eps, alpha, stage, attached=True)
    t_d, t_g, n_d, n_g = 0.0, 0.0, 0, 0
    for i in range(steps):
        t0 = time.perf_counter()
        no.d_step(sd_g, sd_d, z, x_real, eps, alpha, stage, attached=True)
        t1 = time.perf_counter()
        t_d += t1 - t0; n_d += 1
        if i % 5 == 0:
            no.g_step(sd_g, sd_d, z, alpha, stage, critic_grads=True)
            t_g += time.perf_counter() - t1; n_g += 1
        if t_d + t_g > budget_s:
            break
    v = 1.0 / (t_d / n_d + (t_g / n_g) / 5.0)
    return {"value": v, "unit": "steps/s", "cores": th.get_num_threads(), "kind": "port", "steps_run": n_d,
            "sample": f"{n_d} critic step(s) ({t_d / n_d:.2f} s each) + {n_g} generator step(s) ({t_g / n_g:.2f} s each) after {warmup} "
                      f"full-size warm-up(s), {res}x{res}, batch {batch}, fp32, reference's full cost (attached fake batch)"}


def run_reference(args, rank):
    if rank != 0:
        return None
    batch, stage = args.batch, args.stage
    res = 4 * 2 ** stage
    cpu = cpu_baseline(batch, stage=stage, steps=max(1, args.steps), warmup=max(0, args.warmup), budget_s=240.0)
    return {
        "impl": "reference", "metric": "GAN train steps/s", "value": cpu["value"], "unit": "steps/s", "n_gpus": args.gpus,
        "steps": cpu["steps_run"], "warmup": args.warmup, "ms_per_step": 1e3 / cpu["value"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ProGAN WGAN-GP iteration at {res}x{res} (BASELINE config {2 if batch == 8 and stage == 7 else 4}) on the host "
                               "CPU: oracle port of the reference step body, critic step every iteration + generator step every 5th",
                   "stage": stage, "batch_per_gpu": batch, "time_budget_s": 240},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def run_torch_cuda(args, rank, ClockSampler, local):
    """`--impl torch_cuda`: the stock-torch arm as a line of its own (bf16 autocast + channels_last = its best setting)."""
    if rank != 0:
        return None
    from baseline import torch_b200 as tb
    batch, stage = args.batch, args.stage
    res = 4 * 2 ** stage
    with ClockSampler(local) as cs:
        nums = torch_b200_numbers(stage, batch, steps=max(3, args.steps), warmup=max(3, args.warmup))
    best = max((m for m in ("fp32", "bf16_autocast_channels_last") if "value" in nums[m]), key=lambda m: nums[m]["value"])
    return {
        "impl": "torch_cuda", "metric": "GAN train steps/s", "value": nums[best]["value"], "unit": "steps/s", "n_gpus": 1,
        "steps": max(3, args.steps), "warmup": max(3, args.warmup), "ms_per_step": nums[best]["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if best.startswith("bf16") else "f32 (cuDNN TF32 default)", "data": "synthetic",
        "config": {"workload": f"ProGAN WGAN-GP iteration at {res}x{res} through stock torch on the GPU ({best})", "stage": stage, "batch_per_gpu": batch},
        "clocks": cs.summary(), "torch_b200": nums, "gpu_launches": 0,
    }


# ------------------------------------------------------------------------------------------------------------------
# BASELINE config 3: the progressive-growing sweep, every stage at a fixed batch
# ------------------------------------------------------------------------------------------------------------------
def run_sweep(args, rank, world, local, timed_region, ClockSampler, peaks):
    from . import parallel
    from .graphed import GraphedSteps
    dev = th.device("cuda", local)
    batch = args.batch
    pk = peaks()
    rows = []
    th.cuda.manual_seed(1000 + rank)
    with ClockSampler(local) as cs:
        for stage in range(8):
            gen, disc = _build(stage, 0, dev)
            opt_g = th.optim.Adam(gen.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
            opt_d = th.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.0, 0.9), capturable=True, fused=True)
            bucket_d = parallel.FlatGradBucket(disc.parameters()) if world > 1 else None
            bucket_g = parallel.FlatGradBucket(gen.parameters()) if world > 1 else None
            res = 4 * 2 ** stage
            reals = [th.rand(batch, 2, res, res, device=dev) * 2 - 1 for _ in range(4)]
            graphed = GraphedSteps(gen, disc, opt_g, opt_d, batch, 32, res, 0.5, bucket_d=bucket_d, bucket_g=bucket_g)
            it = [0]

            def step():
                graphed.critic_step(reals[it[0] % 4])
                if it[0] % 5 == 0:
                    graphed.generator_step()
                it[0] += 1

            for alpha in ((1.0,) if stage == 0 else (0.5, 1.0)):
                graphed.set_alpha(alpha)
                it[0] = 0
                steps = args.steps if stage >= 5 else 4 * args.steps      # sub-millisecond stages: more steps per timing
                ms = timed_region(step, steps, args.warmup, world)
                sps = world * steps / (ms * 1e-3)
                tf = batch * flop_per_iter(stage) * sps / 1e12
                rows.append({"stage": stage, "resolution": res, "alpha": alpha, "steps_per_s": sps, "ms_per_step": ms / steps,
                             "samples_per_s": sps * batch, "algorithmic_tflops": tf, "frac_of_bf16_peak": tf / (pk["tf_sus"] * world),
                             "regime": "launch / latency bound (whole step < 2 ms)" if ms / steps < 2.0 else "tensor / HBM"})
            del graphed, gen, disc, opt_g, opt_d, reals
            th.cuda.empty_cache()
    if rank != 0:
        return None
    head = next(r for r in rows if r["stage"] == 7 and r["alpha"] == 0.5)
    # CPU: batch 8, every stage once, scaled per sample to the sweep's batch
    cpu_rows, t_total = [], 0.0
    for stage in range(8):
        c = cpu_baseline(8, stage=stage, steps=1, warmup=0)
        cpu_rows.append({"stage": stage, "steps_per_s_at_this_batch": c["value"] * 8.0 / batch, "sample": c["sample"]})
    return {
        "metric": "GAN train steps/s", "value": head["steps_per_s"], "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"progressive-growing sweep (BASELINE config 3): one WGAN-GP iteration (critic + 1/5 generator step, Adam "
                               f"included) at every stage 0..7, batch {batch} per GPU, alpha 0.5 and 1.0; headline value = stage 7, alpha 0.5",
                   "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": f"dp{world}", "cuda_graphs": True,
                   "l2": "stages >= 4: activations of a step exceed L2; stages <= 3 fit in L2 and are launch bound (said per row)"},
        "clocks": cs.summary(), "stages": rows,
        "e2e": None, "gpu_launches": None,
        "roofline": {"bound": "tensor", "kernel": "whole iteration at stage 7 (algorithmic FLOPs of the reference step / time)",
                     "achieved": head["algorithmic_tflops"], "peak": pk["tf_sus"] * world, "unit": "TFLOP/s",
                     "frac": head["frac_of_bf16_peak"], "peak_source": pk["src"] + " (sustained)", "traffic": None},
        "cpu_baseline": {"value": cpu_rows[7]["steps_per_s_at_this_batch"], "unit": "steps/s", "cores": th.get_num_threads(), "kind": "port",
                         "sample": f"batch 8 per stage, scaled per sample to batch {batch}", "stages": cpu_rows},
    }
