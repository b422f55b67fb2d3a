"""CUDA-graph capture of the two optimisation steps (reference train.py:143-175 and :191-214).

At the final resolutions one WGAN-GP iteration is ~250 tensor-core launches plus the pointwise kernels and the
optimiser: issued from Python it is host bound.  The critic step and the generator step are therefore captured
ONCE each (after a warm-up on a side stream) and replayed; all launches of libmusicgan_b200.so go to torch's current
stream, so they are captured like any ATen kernel.  Packed-weight caches are invalidated before capture so that the
weight re-pack kernels are part of the graph (weights change between replays).

The fade-in weight `alpha` is a graph INPUT (a 0-dim device tensor read by the two blend kernels): `set_alpha()` moves
it between replays, so one capture serves a whole stage of the progressive schedule; `train` re-captures at every
`next_layer()` (new modules, new resolution).
"""
from __future__ import annotations

import os

import torch as th

from . import networks
from .train_step import frozen
from .networks import ops


class GraphedSteps:
    def __init__(self, gen, disc, optim_gen, optim_disc, batch: int, rand_channels: int, resolution: int, alpha: float,
                 bucket_d=None, bucket_g=None, warmup: int = 3, static_noise: bool = False,
                 two_streams: bool = None, preserve_state: bool = False):
        self.gen, self.disc, self.og, self.od = gen, disc, optim_gen, optim_disc
        self.batch = batch
        self.alpha = th.full((), float(alpha), dtype=th.float32, device=next(gen.parameters()).device)
        self._alpha_host = float(alpha)
        # parallel.FlatGradBucket of each network (data-parallel runs): the freshly computed gradients are gathered into
        # the bucket, averaged over the ranks by ONE all-reduce node of the graph, and the parameters' .grad become views
        # of the bucket
        self.bucket_d, self.bucket_g = bucket_d, bucket_g
        dev = next(gen.parameters()).device
        self.x_real = th.zeros(batch, 2, resolution, resolution, device=dev)
        self.z_shape = (batch, rand_channels, 2, 2)
        # static_noise: latent vectors and the penalty's uniform sample are graph INPUTS (self.z, self.eps) instead of
        # being drawn inside the graph -- lets a test replay exactly what an eager step computes
        self.z = th.zeros(self.z_shape, device=dev) if static_noise else None
        self.eps = th.zeros(batch, 1, 1, 1, device=dev) if static_noise else None
        self.d_stats = self.g_stats = None
        self._gd = self._gg = None
        # two_streams: the penalty branch (forward on x_hat, inner backward, double backward) and the Wasserstein branch
        # (forward / backward on the real + fake batch) of the critic step only share read-only inputs, so they are
        # captured as parallel branches of the graph: most of their ~600 launches are small-spatial layers that use a
        # fraction of the SMs and now overlap pairwise
        if two_streams is None:
            two_streams = os.environ.get("MG_TWO_STREAMS", "1") != "0"
        self._branch = th.cuda.Stream(device=dev) if two_streams else None
        self._wgrad_lanes = os.environ.get("MG_WGRAD_LANES", "1") != "0"
        if two_streams and hasattr(th.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            # the parameters' AccumulateGrad nodes were created on another stream; gradients are collected by
            # autograd.grad (never accumulated), so the mismatch the engine warns about is intended
            th.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        # preserve_state: the warm-up iterations before the capture take real optimiser steps (on whatever sits in the
        # static input buffer); `train` must not see them -- parameters and optimiser state are put back afterwards
        saved = self._snapshot() if preserve_state else None
        self._capture(warmup)
        if saved is not None:
            self._restore(saved)

    def set_alpha(self, alpha: float) -> None:
        """New fade-in weight for the following replays (one asynchronous scalar write on the current stream)."""
        if alpha != self._alpha_host:
            self.alpha.fill_(float(alpha))
            self._alpha_host = float(alpha)

    # -- the two step bodies (static shapes, no host sync) -------------------------------------------------
    def _critic_body(self):
        gen, disc, alpha, n = self.gen, self.disc, self.alpha, self.batch
        z = self.z if self.z is not None else th.randn(self.z_shape, device=self.x_real.device)
        ops.prepack(gen)                           # one launch re-packs every weight of a network after its last update
        ops.prepack(disc)                          # (from here on both branches below only read the packed copies)
        with th.no_grad():
            x_fake = gen(z, alpha)
        params = [p for p in disc.parameters()]
        if self._branch is None:
            with self._lane(params) as lane:
                out = disc(th.cat([self.x_real, x_fake], dim=0), alpha)
                d_loss = networks.wasserstein_discriminator_loss(out[:n], out[n:])
                gp = disc.gradient_penalty(self.x_real, x_fake, alpha, eps=self.eps)
                grads = lane.merge(th.autograd.grad(d_loss + gp, params, allow_unused=True))
        else:
            main, branch = th.cuda.current_stream(), self._branch
            branch.wait_stream(main)
            with th.cuda.stream(branch), self._lane(params) as lane:
                gp = disc.gradient_penalty(self.x_real, x_fake, alpha, eps=self.eps)
                grads_gp = lane.merge(th.autograd.grad(gp, params, allow_unused=True))
            with self._lane(params) as lane:
                out = disc(th.cat([self.x_real, x_fake], dim=0), alpha)
                d_loss = networks.wasserstein_discriminator_loss(out[:n], out[n:])
                grads_w = lane.merge(th.autograd.grad(d_loss, params, allow_unused=True))
            main.wait_stream(branch)
            # sum of the two branches' gradients: one multi-tensor add instead of one tiny kernel per parameter
            both = [(gw, gg) for gw, gg in zip(grads_w, grads_gp) if gw is not None and gg is not None]
            if both:
                th._foreach_add_([gw for gw, _ in both], [gg for _, gg in both])
            grads = [gw if gw is not None else gg for gw, gg in zip(grads_w, grads_gp)]
        self._install(params, grads, self.bucket_d)
        self.od.step()
        return th.stack([d_loss.detach(), gp.detach(), out[:n].mean().detach(), out[n:].mean().detach()])

    def _generator_body(self):
        gen, disc, alpha = self.gen, self.disc, self.alpha
        z = self.z if self.z is not None else th.randn(self.z_shape, device=self.x_real.device)
        # only G's gradients are needed (the reference computes and discards D's, train.py:208-214): with the critic
        # frozen its weight-gradient kernels are not even launched
        params = [p for p in gen.parameters()]
        ops.prepack(gen)
        ops.prepack(disc)
        with frozen(disc), self._lane(params) as lane:
            out_fake = disc(gen(z, alpha), alpha)
            g_loss = networks.wasserstein_generator_loss(out_fake)
            grads = lane.merge(th.autograd.grad(g_loss, params, allow_unused=True))
        self._install(params, grads, self.bucket_g)
        self.og.step()
        return th.stack([g_loss.detach(), out_fake.mean().detach()])

    def _lane(self, params):
        # weight gradients on a second stream beside the data-gradient chain (ops.WgradLane); depth 0 = in line
        return ops.WgradLane(params if self._wgrad_lanes else [])

    @staticmethod
    def _install(params, grads, bucket):
        if bucket is None:
            for p, g in zip(params, grads):
                p.grad = g
        else:
            assert [id(p) for p in bucket.params] == [id(p) for p in params]
            bucket.adopt(grads)

    def _snapshot(self):
        params = [p for m in (self.gen, self.disc) for p in m.parameters()]
        state = {}
        for opt in (self.og, self.od):
            for p, st in opt.state.items():
                state[id(p)] = {k: (v.clone() if th.is_tensor(v) else v) for k, v in st.items()}
        return [p.detach().clone() for p in params], state

    def _restore(self, saved):
        values, state = saved
        params = [p for m in (self.gen, self.disc) for p in m.parameters()]
        with th.no_grad():
            for p, v in zip(params, values):
                p.copy_(v)                           # in place: the graphs hold the storages
            for opt in (self.og, self.od):
                for p, st in opt.state.items():
                    old = state.get(id(p))
                    for k, v in st.items():
                        if th.is_tensor(v):
                            v.copy_(old[k]) if old is not None else v.zero_()
        ops.invalidate_pack_cache()

    def _capture(self, warmup: int):
        side = th.cuda.Stream()
        side.wait_stream(th.cuda.current_stream())
        with th.cuda.stream(side):
            for _ in range(warmup):
                self._critic_body()
                self._generator_body()
        th.cuda.current_stream().wait_stream(side)
        th.cuda.synchronize()
        # the job tables of the one-launch weight packing are uploaded here, outside the capture (a host-to-device copy
        # of pageable memory cannot be captured); inside the graphs prepack() only launches the kernel
        ops.prepack(self.disc)
        ops.prepack(self.gen)
        th.cuda.synchronize()
        ops.invalidate_pack_cache()
        # kernel scratch requested during the captures comes from the graphs' private pool: it is kept in a dict that
        # lives and dies with this object (ops.capture_workspaces), not in the process-wide workspace cache
        self._ws = {}
        self._gd = th.cuda.CUDAGraph()
        with ops.capture_workspaces(self._ws), th.cuda.graph(self._gd):
            self.d_stats = self._critic_body()
        ops.invalidate_pack_cache()
        self._gg = th.cuda.CUDAGraph()
        with ops.capture_workspaces(self._ws), th.cuda.graph(self._gg, pool=self._gd.pool()):
            self.g_stats = self._generator_body()
        ops.invalidate_pack_cache()

    # -- replay ---------------------------------------------------------------------------------------------
    def critic_step(self, x_real: th.Tensor) -> th.Tensor:
        self.x_real.copy_(x_real, non_blocking=True)
        self._gd.replay()
        ops.invalidate_pack_cache()      # the optimiser step inside the graph fires no hook: packed weights are stale now
        return self.d_stats          # [d_loss, grad_pen, mean D(real), mean D(fake)] (device tensor, overwritten on replay)

    def generator_step(self) -> th.Tensor:
        self._gg.replay()
        ops.invalidate_pack_cache()
        return self.g_stats          # [g_loss, mean D(fake)]
