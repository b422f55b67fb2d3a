#!/usr/bin/env python
"""bench.py -- throughput of the MusicGAN hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload both|transform|train|sweep|generate]
                    [--batch B] [--stage S] [--nb-vec V] [--impl b200|reference|torch_cuda]

Prints ONE JSON line (rank 0).  Workloads (BASELINE.json configs):
  transform  config 1 scaled to a batch: `--clips` 60 s mono clips per GPU resident in HBM; a step is
             one pass of mg_stft_magif_f32 (STFT -> magnitude/IF chunks) over the batch; frames/s.
  train      config 2 (HEADLINE): one WGAN-GP iteration of the reference schedule at 512x512, batch 8 per GPU:
             a critic step every iteration, a generator step every 5th (train.py:189), Adam updates included; steps/s.
  both       default: the train line, with the transform line nested under "secondary".
  sweep      config 3: the same iteration at every stage 0..7 of the progressive schedule (`--batch 64`), alpha 0.5 / 1.0.
  train --batch 64 under torchrun = config 4 (64 samples per GPU, one flat-bucket all-reduce per optimiser step; `--stage 3`
             for the 32 x 32 stage).
  generate   config 5: G inference on a wide latent + inverse transform, `--gen-clips` clips per GPU, `--nb-vec` 4 | 10.
`--impl torch_cuda` times the reference's modules through stock torch (cuDNN / cuFFT) on the same GPU: the bar on the box.

`value` is timed with inputs resident in HBM; `e2e` goes through the public Python API with pinned-host
inputs and a device->host read of the step's result inside the timed region.  `--impl reference` times the
CPU oracle port of the reference (all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

BYTES_PER_FRAME = 5120           # 256*4 B in + 2*512*4 B out (same figure for the inverse)
N_60S = 2_646_000


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def timed_region(step_fn, steps: int, warmup: int, world: int):
    """W warm-up steps, then exactly K steps between barrier+synchronize, CUDA events, max over ranks."""
    import torch.distributed as dist
    for _ in range(warmup):
        step_fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step_fn()
    b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


# ------------------------------------------------------------------------------------------------
# transform workload
# ------------------------------------------------------------------------------------------------
def cpu_transform_baseline(seconds: float = 10.0):
    """Oracle port of the reference's wav_to_stft + stft_to_phase_magn on the host cores, 60 s clips."""
    from oracle import audio_oracle as ao, cases
    torch.set_num_threads(os.cpu_count() or 1)
    wav = cases.forward_wav("noise_60s")[0]
    ao.wav_to_magn_phase(wav)                      # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        ao.wav_to_magn_phase(wav)
        n += 1
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    frames = n * (1 + N_60S // 256)
    return frames / el, n, el


def synthetic_wavs(batch: int, n: int, seed: int) -> torch.Tensor:
    """(batch, n) mono uniform-noise clips in [-0.5, 0.5) (the product arm imports nothing from oracle/)."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(batch, n, generator=g) * 2 - 1) * 0.5


def run_transform(args, rank, world, local):
    from musicgan_b200 import _lib, audio
    clips = args.clips
    plan = audio.ForwardPlan(N_60S, clips, 1)
    frames_per_step = clips * plan.T
    host = synthetic_wavs(clips, N_60S, seed=2024 + rank).pin_memory()
    dev = host.cuda(non_blocking=True)
    torch.cuda.synchronize()

    def step():
        plan.run(dev)

    _lib.profile_enable(True)
    with ClockSampler(local) as cs:
        ms = timed_region(step, args.steps, args.warmup, world)
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    value = world * frames_per_step * args.steps / (ms * 1e-3)

    # end to end through the public API: pinned host wav -> H2D -> transform -> D2H of the chunks
    e2e_clips = min(clips, args.e2e_clips)
    plan2 = audio.ForwardPlan(N_60S, e2e_clips, 1) if e2e_clips != clips else plan
    host_in = host[:e2e_clips].contiguous().pin_memory()
    out_m = torch.empty(plan2.magn.shape, dtype=torch.float32).pin_memory()
    out_p = torch.empty(plan2.phase.shape, dtype=torch.float32).pin_memory()
    dev_in = torch.empty_like(host_in, device="cuda")

    def e2e_step():
        dev_in.copy_(host_in, non_blocking=True)
        m, p = audio.wav_to_magn_phase_batch(dev_in, plan2)
        out_m.copy_(m, non_blocking=True)
        out_p.copy_(p, non_blocking=True)

    e2e_steps = max(1, min(args.steps, 5))
    ms2 = timed_region(e2e_step, e2e_steps, 1, world)
    e2e_value = world * e2e_clips * plan2.T * e2e_steps / (ms2 * 1e-3)

    res = None
    if rank == 0:
        pk = peaks()
        k_ms, k_n = prof.get("k_stft", (0.0, 0))
        per_launch_s = (k_ms / max(k_n, 1)) * 1e-3
        achieved = frames_per_step * BYTES_PER_FRAME / per_launch_s / 1e9 if per_launch_s > 0 else 0.0
        step_gbs = frames_per_step * BYTES_PER_FRAME * args.steps / (ms * 1e-3) / 1e9
        cpu_v, cpu_n, cpu_el = cpu_transform_baseline(args.cpu_seconds)
        torch_arm = None
        if world == 1:
            from baseline import torch_b200 as tb
            torch.cuda.empty_cache()
            try:
                tv, tms = tb.time_transform(16, N_60S, 3, 2)
                torch_arm = {"value": tv, "unit": "frames/s", "ms_per_step": tms, "clips_per_step": 16,
                             "what": "torch-CUDA restatement of wav_to_stft + stft_to_phase_magn (cuFFT + ATen, cumsum in double), "
                                     "batched over 16 clips, inputs resident"}
            except Exception as e:      # noqa: BLE001
                torch_arm = {"error": f"{type(e).__name__}: {str(e)[:120]}"}
        res = {
            "metric": "spectrogram frames/s (STFT+IF)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "create_dataset transform (BASELINE config 1): 60 s mono 44.1 kHz clips, n_fft 1024, hop 256, "
                                   "nb_vec 512 -> 20 chunks/clip", "clips_per_gpu": clips, "frames_per_clip": plan.T,
                       "l2": f"inputs larger than L2 ({clips * N_60S * 4 / 1e6:.0f} MB wav + {plan.ws_bytes / 1e6:.0f} MB scratch per step)",
                       "parallelism": f"clips sharded over {world} GPU(s), no collective"},
            "clocks": cs.summary(),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": e2e_clips * N_60S * 4,
                    "d2h_bytes_per_step": int(out_m.numel() + out_p.numel()) * 4, "clips_per_step": e2e_clips},
            "gpu_launches": int(sum(c for _, c in prof.values())),
            "roofline": {"bound": "hbm", "kernel": "k_stft", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": achieved / pk["hbm"], "peak_source": pk["src"],
                         # dram__bytes_read + write of one k_stft launch (ncu --set full, profiles/r02_prof_stft_raw.csv:
                         # 366.2 MB for 8 x 10 336 frames = 4 429 B / frame against 5 120 algorithmic: the tail of the
                         # 4 KB / frame scratch is still in L2 when the launch ends), scaled to this launch's frames
                         "traffic": 4429.0 * frames_per_step, "traffic_source": "ncu, bytes per frame x frames of this launch",
                         "step_achieved": step_gbs, "step_frac": step_gbs / pk["hbm"],
                         "kernel_ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items()}},
            "cpu_baseline": {"value": cpu_v, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{cpu_n} x 60 s clip (oracle port of wav_to_stft + stft_to_phase_magn) in {cpu_el:.1f} s"},
            "torch_b200": torch_arm,
        }
    return res


def run_transform_reference(args, rank):
    if rank != 0:
        return None
    # each "step" = one 60 s clip through the CPU oracle port of the reference
    from oracle import audio_oracle as ao, cases
    torch.set_num_threads(os.cpu_count() or 1)
    wav = cases.forward_wav("noise_60s")[0]
    for _ in range(max(1, min(args.warmup, 2))):
        ao.wav_to_magn_phase(wav)
    steps = min(args.steps, 40)
    t0 = time.perf_counter()
    for _ in range(steps):
        ao.wav_to_magn_phase(wav)
    el = time.perf_counter() - t0
    v = steps * (1 + N_60S // 256) / el
    return {
        "impl": "reference", "metric": "spectrogram frames/s (STFT+IF)", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": el / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "create_dataset transform (BASELINE config 1): 60 s mono clips; one clip per step on the host CPU"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{steps} x 60 s clip"},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


# ------------------------------------------------------------------------------------------------
# generate workload (BASELINE config 5): G inference on a wide latent + inverse transform, clips sharded over the ranks
# ------------------------------------------------------------------------------------------------
def cpu_generate_baseline(nb_vec: int, n_clips: int = 4):
    """Oracle port of generate.py:47-65 on the host cores: G forward on the wide latent (fp32) + per-clip inverse transform."""
    from oracle import audio_oracle as ao, networks_oracle as no
    torch.set_num_threads(os.cpu_count() or 1)
    sd_g = no.make_state("gen", 7, 1)
    g = torch.Generator().manual_seed(0)
    z = torch.randn(n_clips, 32, 2, 2 * nb_vec, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        img = torch.cat([no.gen_forward(sd_g, z[i:i + 1], 1.0, 7) for i in range(n_clips)])
    t1 = time.perf_counter()
    for i in range(n_clips):
        ao.magn_phase_to_wav(img[i:i + 1].clone())
    t2 = time.perf_counter()
    return {"value": n_clips / (t2 - t0), "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_clips} clips of {512 * nb_vec} frames: G forward {t1 - t0:.1f} s + inverse transform {t2 - t1:.1f} s"}


def run_generate(args, rank, world, local):
    from musicgan_b200 import _lib, audio, networks
    nb_vec, per_gpu = args.nb_vec, args.gen_clips
    sub = min(per_gpu, max(1, 128 // nb_vec))          # 128 x 512 frame-columns per sub-batch: >= 148 CTAs in every inverse kernel
    torch.manual_seed(0)
    gen = networks.Generator(32, end_layer=7).eval().cuda()
    z_all = torch.randn(per_gpu, 32, 2, 2 * nb_vec, device="cuda")
    host_out = torch.empty(sub, 256 * (512 * nb_vec - 1)).pin_memory()

    def step(e2e=False):
        with torch.no_grad():
            for lo in range(0, per_gpu, sub):
                img = gen(z_all[lo:lo + sub], 1.0)
                wav = audio.magn_phase_to_wave_batch(img, 1)
                if e2e:
                    host_out[: wav.size(0)].copy_(wav, non_blocking=True)

    _lib.profile_enable(True)
    with ClockSampler(local) as cs:
        ms = timed_region(step, args.steps, args.warmup, world)
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    value = world * per_gpu * args.steps / (ms * 1e-3)
    ms2 = timed_region(lambda: step(True), max(1, min(args.steps, 5)), 1, world)
    e2e = world * per_gpu * max(1, min(args.steps, 5)) / (ms2 * 1e-3)
    if rank != 0:
        return None
    pk = peaks()
    frames = per_gpu * 512 * nb_vec * args.steps
    inv_ms = sum(v[0] for k, v in prof.items() if k.startswith("k_inv") or k == "k_istft")
    achieved = frames * BYTES_PER_FRAME / (inv_ms * 1e-3) / 1e9 if inv_ms > 0 else 0.0
    return {
        "metric": "generated clips/s (G inference + IF->phase->iSTFT)", "value": value, "unit": "clips/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16 (G) / f32 (inverse)", "data": "synthetic",
        "config": {"workload": f"generate (BASELINE config 5): {per_gpu} clips/GPU/step of {512 * nb_vec} frames "
                               f"({256 * (512 * nb_vec - 1) / 44100:.1f} s), nb_vec {nb_vec}, sub-batches of {sub}",
                   "clips_total": per_gpu * world, "parallelism": f"clips sharded over {world} GPU(s), no collective"},
        "clocks": cs.summary(),
        "e2e": {"value": e2e, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": per_gpu * 256 * (512 * nb_vec - 1) * 4},
        "gpu_launches": int(sum(c for _, c in prof.values())),
        "roofline": {"bound": "hbm", "kernel": "inverse transform kernels (k_inv_* + k_istft)", "achieved": achieved, "peak": pk["hbm"],
                     "unit": "GB/s", "frac": achieved / pk["hbm"], "peak_source": pk["src"], "traffic": None,
                     "kernel_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}},
        "cpu_baseline": cpu_generate_baseline(nb_vec, 4),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch_cuda"])
    ap.add_argument("--workload", default=os.environ.get("MG_BENCH_WORKLOAD", "both"), choices=["both", "transform", "train", "generate", "sweep"],
                    help="both (default): the train workload (BASELINE config 2) is the headline line, the transform workload "
                         "(config 1) rides along under 'secondary'")
    ap.add_argument("--clips", type=int, default=64, help="clips per GPU per step (transform workload)")
    ap.add_argument("--batch", type=int, default=8, help="samples per GPU per step (train / sweep workloads; 64 = BASELINE configs 3 and 4)")
    ap.add_argument("--stage", type=int, default=7, help="growth stage of the train workload (7 = 512 x 512)")
    ap.add_argument("--nb-vec", type=int, default=4, help="512-frame images per generated clip (generate workload; CLI default of the reference: 10)")
    ap.add_argument("--e2e-clips", type=int, default=16)
    ap.add_argument("--gen-clips", type=int, default=128, help="clips per GPU per step (generate workload)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first
    # communicator): everything goes to stderr while the bench runs, the result line goes to the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    if args.impl == "torch_cuda":
        rank = int(os.environ.get("RANK", "0"))
        if rank == 0:
            torch.cuda.set_device(0)
            from musicgan_b200 import bench_train
            res = bench_train.run_torch_cuda(args, rank, ClockSampler, 0)
            if args.workload in ("both", "transform"):
                from baseline import torch_b200 as tb
                tv, tms = tb.time_transform(16, N_60S, max(3, min(args.steps, 10)), 2)
                res["secondary"] = {"impl": "torch_cuda", "metric": "spectrogram frames/s (STFT+IF)", "value": tv, "unit": "frames/s",
                                    "ms_per_step": tms, "config": {"workload": "create_dataset transform, 16 x 60 s clips per step, stock torch on the GPU"}}
            emit(res)
        return
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        res = None
        if args.workload in ("both", "train"):
            from musicgan_b200 import bench_train
            res = bench_train.run_reference(args, rank)
        if args.workload in ("both", "transform"):
            sec = run_transform_reference(args, rank)
            if res is None:
                res = sec
            elif sec is not None:
                res["secondary"] = sec
        if res is not None:
            emit(res)
        return

    rank, world, local = dist_setup(args.gpus)
    res = None
    if args.workload == "generate":
        res = run_generate(args, rank, world, local)
    if args.workload == "sweep":
        from musicgan_b200 import bench_train
        res = bench_train.run_sweep(args, rank, world, local, timed_region, ClockSampler, peaks)
    if args.workload in ("both", "train"):
        from musicgan_b200 import bench_train
        res = bench_train.run(args, rank, world, local, timed_region, ClockSampler, peaks)
        torch.cuda.empty_cache()
    if args.workload in ("both", "transform"):
        sec = run_transform(args, rank, world, local)
        if args.workload == "transform":
            res = sec
        elif res is not None and sec is not None:
            res["secondary"] = sec
    if rank == 0 and res is not None:
        emit(res)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
