"""Raw (non-autograd) wrappers of the convolution entry points of the C ABI.

Activations are logical (B, C, H, W) bf16 tensors in torch.channels_last memory format, i.e. NHWC in
memory -- exactly what the kernels read and write.  Weights are the fp32 master tensors
(Cout, Cin, 3, 3) of the drop-in nn.Modules (same shapes as the reference's nn.Conv2d)."""
from __future__ import annotations

import collections
import ctypes
import os
from ctypes import c_int, c_size_t, c_void_p

import torch as th

from .. import _lib

FLAG_LRELU, FLAG_PIXELNORM, FLAG_UPSAMPLE_IN, FLAG_DGRAD, FLAG_SPLIT_W, FLAG_W3 = 1, 2, 4, 8, 16, 32

# Layers whose OUTPUT height is <= PRECISE_MAX_RES run on fp32 activations with split-bf16 operands (conv_split.cu): on
# those few-pixel layers bf16 operand rounding flips LeakyReLU masks and moves the WGAN-GP gradients by 2-10 %
# (scripts/precision_study.py; north_star asks for 1e-2).  With 32 the batch-1 golden cases at stages 6 / 7 sit at
# 1.0-1.1e-2, with 64 at 7-8e-3 (emulated and measured).  0 switches the precise path off.
PRECISE_MAX_RES = int(os.environ.get("MG_PRECISE_MAX_RES", "64"))
# Forward convolutions of the bf16 layers take their weights as hi + lo bf16 pairs (flag 16); 0 = plain bf16 weights.
SPLIT_W = os.environ.get("MG_SPLIT_W", "1") != "0"
# ... up to this output height.  Measured on B200: leaving the 512 x 512 layers on plain bf16 weights (256) buys 1 % of step
# time and moves the stage-7 batch-1 generator-step gradient from 7.8e-3 to 9.1e-3 of the oracle's: not taken.
SPLIT_W_MAX_RES = int(os.environ.get("MG_SPLIT_W_MAX_RES", "512"))
# The critic's forward convolutions on the precise path take the fp32 weight exactly (three bf16 parts) up to this output
# height: it is the 1 x 1 ... 32 x 32 layers where a single pre-activation within 1e-5 of zero moves the penalty gradient by
# percents (golden case stage2_b3); at 64 x 64 two parts give the same errors to three digits (scripts/precision_study.py
# emulation, stages 4-7: G-step 7.6e-4 / 4.4e-3 / 8.0e-3 / 6.3e-3 either way) and the layer runs 10-40 % faster.
EXACT_W_MAX_RES = int(os.environ.get("MG_EXACT_W_MAX_RES", "32"))


# Inference (module in eval mode, autograd off: `generate`) has no gradients whose masks need protecting: the forward-only
# mode keeps the fp32 split-operand path where it is nearly free (output height <= 32) and runs the rest on plain bf16
# weights.  G output rel-L2 vs the fp32 oracle: see tests/test_networks_gpu.py::test_generator_inference_mode.
FORWARD_ONLY_PRECISE_MAX_RES = int(os.environ.get("MG_INFER_PRECISE_MAX_RES", "32"))
_forward_only = [False]


class forward_only:
    def __init__(self, on: bool):
        self.on = bool(on)

    def __enter__(self):
        self.prev, _forward_only[0] = _forward_only[0], self.on

    def __exit__(self, *exc):
        _forward_only[0] = self.prev
        return False


def is_precise(h_out: int) -> bool:
    return h_out <= (FORWARD_ONLY_PRECISE_MAX_RES if _forward_only[0] else PRECISE_MAX_RES)

class PackJob(ctypes.Structure):        # mgPackJob of include/musicgan_b200.h
    _fields_ = [("w", c_void_p), ("out", c_void_p), ("cout_fwd", c_int), ("cin_fwd", c_int), ("flip", c_int), ("nt", c_int),
                ("parts", c_int), ("kind", c_int), ("total", c_int), ("reserved", c_int)]


_declared = False
_ws_cache = {}
# bench bookkeeping of the 3x3 convolution launches issued through this module: FLOPs (2 * B*H*W * 9*Cin*Cout per call),
# algorithmic HBM bytes (every operand read once, the result written once; weights excluded) and, when "launches" is a
# list, one (flops, bytes) pair per launch so a per-launch roofline can be summed
FLOPS = {"count": 0.0, "bytes": 0.0, "launches": None}


def _account(flops, nbytes):
    FLOPS["count"] += flops
    FLOPS["bytes"] += nbytes
    if FLOPS["launches"] is not None:
        FLOPS["launches"].append((flops, nbytes))


def _l():
    global _declared
    l = _lib.lib()
    if not _declared:
        l.mg_conv3x3_workspace_bytes.restype = c_size_t
        l.mg_conv3x3_workspace_bytes.argtypes = [c_int, c_int]
        l.mg_conv3x3_bf16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]
        if hasattr(l, "mg_conv3x3_wgrad_bf16"):
            l.mg_conv3x3_wgrad_workspace_bytes.restype = c_size_t
            l.mg_conv3x3_wgrad_workspace_bytes.argtypes = [c_int] * 5
            l.mg_conv3x3_wgrad_bf16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                                c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
            l.mg_conv3x3_wgrad_bias_bf16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                                     c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
        c_int64 = ctypes.c_int64
        l.mg_rgb_expand_bf16.argtypes = [c_void_p] * 5 + [c_int, c_int64, c_int, c_int, c_void_p]
        l.mg_rgb_project_bf16.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p]
        l.mg_rgb_wgrad_bf16.argtypes = [c_void_p] * 5 + [c_int, c_int64, c_int, c_void_p]
        l.mg_pool2_bf16.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
        l.mg_colsum_workspace_bytes.restype = c_size_t
        l.mg_colsum_workspace_bytes.argtypes = [c_int]
        l.mg_pixelnorm_lrelu_bwd_bf16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int64, c_int, c_void_p]
        l.mg_lrelu_bwd_bf16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int64, c_int, c_void_p]
        l.mg_unpool2_lrelu_bwd_bf16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_void_p]
        l.mg_pool2_planes_f32.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]
        l.mg_conv3x3_pack_weights.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]
        l.mg_pack_job_fill.argtypes = [ctypes.POINTER(PackJob), c_void_p, c_void_p, c_int, c_int, c_int, c_int]
        l.mg_pack_weights_multi.argtypes = [c_void_p, c_int, c_int, c_void_p]
        l.mg_conv3x3_split_workspace_bytes.restype = c_size_t
        l.mg_conv3x3_split_workspace_bytes.argtypes = [c_int, c_int]
        l.mg_conv3x3_split_pack_weights.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]
        l.mg_conv3x3_split_f32.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 6 + [c_void_p]
        for sfx in ("_f32",):
            getattr(l, "mg_rgb_expand" + sfx).argtypes = l.mg_rgb_expand_bf16.argtypes
            getattr(l, "mg_rgb_project" + sfx).argtypes = l.mg_rgb_project_bf16.argtypes
            getattr(l, "mg_rgb_wgrad" + sfx).argtypes = l.mg_rgb_wgrad_bf16.argtypes
            getattr(l, "mg_pool2" + sfx).argtypes = l.mg_pool2_bf16.argtypes
            getattr(l, "mg_lrelu_bwd" + sfx).argtypes = l.mg_lrelu_bwd_bf16.argtypes
            getattr(l, "mg_unpool2_lrelu_bwd" + sfx).argtypes = l.mg_unpool2_lrelu_bwd_bf16.argtypes
            getattr(l, "mg_pixelnorm_lrelu_bwd" + sfx).argtypes = l.mg_pixelnorm_lrelu_bwd_bf16.argtypes
        _declared = True
    return l


# Scratch buffers are cached per (purpose, device, stream).  A buffer first requested DURING a CUDA-graph capture is carved
# from that graph's private memory pool and dies with the graph: such buffers live in a dict owned by the capturing object
# (graphed.GraphedSteps installs it with `capture_workspaces`), never in the process-wide cache -- a later graph or eager
# call must not be handed an address inside a destroyed graph's pool.
_capture_ws = [None]


class capture_workspaces:
    def __init__(self, owner_dict):
        self.d = owner_dict

    def __enter__(self):
        self.prev, _capture_ws[0] = _capture_ws[0], self.d

    def __exit__(self, *exc):
        _capture_ws[0] = self.prev
        return False


def _workspace(dev, nbytes: int, tag: str = "pack") -> th.Tensor:
    key = (tag, dev.index, th.cuda.current_stream(dev).cuda_stream)
    cache = _ws_cache
    if th.cuda.is_current_stream_capturing():
        if _capture_ws[0] is None:
            raise RuntimeError("musicgan_b200: a kernel workspace was requested inside a CUDA-graph capture outside "
                               "ops.capture_workspaces(...) (use graphed.GraphedSteps)")
        cache = _capture_ws[0]
    ws = cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = th.empty(max(nbytes, 1 << 20), dtype=th.uint8, device=dev)
        cache[key] = ws
    return ws


_pack_epoch = [0]


def invalidate_pack_cache() -> None:
    """Forget every cached packed weight (used around CUDA-graph capture so that the pack kernels are captured)."""
    _pack_epoch[0] += 1


# Fused optimisers (torch._fused_adam_, used by train.py) update parameters WITHOUT bumping Tensor._version, so the
# version stamp alone would keep serving the packed copy of the previous weights.  Every optimiser step of the process
# therefore starts a new cache epoch.
try:
    from torch.optim.optimizer import register_optimizer_step_post_hook
    register_optimizer_step_post_hook(lambda optimizer, args, kwargs: invalidate_pack_cache())
except ImportError:      # pragma: no cover  (older torch: train_step invalidates explicitly)
    pass


def _packed_weights(w: th.Tensor, cin: int, cout: int, kind):
    """Packed copy of a weight tensor, cached ON the parameter object per (version, storage address): parameters change
    once per optimiser step but are read by several forward / backward kernels.  `kind` = ("bf16", mode) with the mode
    bits of mg_conv3x3_pack_weights (1 dgrad, 2 hi + lo, 4 PixelNorm layer) or ("split", dgrad) for conv_split.cu;
    cin / cout are those of the GEMM that will run.  The cache lives and dies with the Parameter (a global dict keyed by
    address would hand one model's weights to the next model allocated at the same address).  Temporaries
    (double-backward operands) are never cached."""
    # parameters only (temporaries change every call); a parameter frozen for one step (train_step.frozen: the critic
    # inside the generator step) is still a parameter
    if os.environ.get("MG_NO_PACK_CACHE") or not (w.is_leaf and (w.requires_grad or getattr(w, "_mg_frozen", False))):
        return None
    cache = getattr(w, "_mg_packed", None)
    if cache is None:
        cache = {}
        w._mg_packed = cache
    stamp = (w._version, w.data_ptr(), _pack_epoch[0])
    key = (kind, cin, cout)
    hit = cache.get(key)
    if hit is not None and hit[0] == stamp:
        return hit[1]
    l = _l()
    split = kind[0] == "split"
    nbytes = l.mg_conv3x3_split_workspace_bytes(cin, cout) if split else l.mg_conv3x3_workspace_bytes(cin, cout)
    if hit is not None and hit[1].device == w.device:
        buf = hit[1]
    else:
        if th.cuda.is_current_stream_capturing():      # would live in (and die with) the capturing graph's memory pool
            raise RuntimeError("musicgan_b200: a packed-weight buffer was first requested inside a CUDA-graph capture; run "
                               "the step once eagerly before capturing (graphed.GraphedSteps warms up)")
        buf = th.empty(nbytes, dtype=th.uint8, device=w.device)
    with th.cuda.device(w.device):
        if split:
            _lib.check(l.mg_conv3x3_split_pack_weights(w.data_ptr(), cin, cout, kind[1], buf.data_ptr(), buf.numel(),
                                                       th.cuda.current_stream().cuda_stream), "mg_conv3x3_split_pack_weights")
        else:
            _lib.check(l.mg_conv3x3_pack_weights(w.data_ptr(), cin, cout, kind[1], buf.data_ptr(), buf.numel(),
                                                 th.cuda.current_stream().cuda_stream), "mg_conv3x3_pack_weights")
    cache[key] = (stamp, buf)
    return buf


_multi_tables = {}


def prepack(module) -> None:
    """Refresh, on the current stream and in ONE launch, every packed copy the parameters of `module` have been asked for so
    far (the warm-up steps before a graph capture ask for all of them).  After this call the forward / data-gradient kernels
    only READ the cached copies, so independent branches of a step may run on different streams (graphed.py)."""
    entries = []
    for p in module.parameters():
        cache = getattr(p, "_mg_packed", None)
        if cache and p.is_cuda:
            for key, (stamp, buf) in cache.items():
                entries.append((p, key, buf))
    if not entries:
        return
    l = _l()
    dev = entries[0][0].device
    ident = tuple((p.data_ptr(), buf.data_ptr(), key) for p, key, buf in entries)
    hit = _multi_tables.get(id(module))
    if hit is None or hit[0] != ident:
        jobs = (PackJob * len(entries))()
        for j, (p, (kind, cin, cout), buf) in zip(jobs, entries):
            _lib.check(l.mg_pack_job_fill(ctypes.byref(j), p.data_ptr(), buf.data_ptr(), cin, cout, 1 if kind[0] == "split" else 0, kind[1]),
                       "mg_pack_job_fill")
        table = th.frombuffer(bytearray(bytes(jobs)), dtype=th.uint8).to(dev)
        hit = (ident, table, len(entries), max(j.total for j in jobs))
        _multi_tables[id(module)] = hit
    _, table, n, max_total = hit
    with th.cuda.device(dev):
        _lib.check(l.mg_pack_weights_multi(table.data_ptr(), n, max_total, th.cuda.current_stream().cuda_stream), "mg_pack_weights_multi")
    for p, key, buf in entries:
        p._mg_packed[key] = ((p._version, p.data_ptr(), _pack_epoch[0]), buf)


def _check_act(x: th.Tensor, name: str, dtype=None):
    if not (x.is_cuda and x.dtype in (th.bfloat16, th.float32) and x.dim() == 4) or (dtype is not None and x.dtype != dtype):
        raise TypeError(f"{name}: expected a CUDA {dtype or 'bf16 / fp32'} (B, C, H, W) tensor, got {x.dtype} {tuple(x.shape)} on {x.device}")
    if not x.is_contiguous(memory_format=th.channels_last):
        raise ValueError(f"{name}: activations must be channels_last (NHWC in memory)")


def as_act(x: th.Tensor, dtype=th.bfloat16) -> th.Tensor:
    """`dtype` (bf16, or fp32 on the precise path) + channels_last (no copy if already so)."""
    return x.to(dtype=dtype).contiguous(memory_format=th.channels_last)


def keep_act(x: th.Tensor) -> th.Tensor:
    """channels_last in the tensor's own activation dtype (fp32 stays fp32, everything else becomes bf16)."""
    return as_act(x, th.float32 if x.dtype == th.float32 else th.bfloat16)


def _sfx(x: th.Tensor) -> str:
    return "_f32" if x.dtype == th.float32 else "_bf16"


def conv3x3(x: th.Tensor, w: th.Tensor, bias=None, *, lrelu=False, pixelnorm=False, upsample_in=False,
            dgrad=False, want_inv_norm=False, split_w=False, exact_w=False):
    """y = conv3x3(x) (+bias) (+LeakyReLU 0.2) (+PixelNorm); `upsample_in` reads x through a nearest x2
    upsampling; `dgrad` computes the data gradient of the forward conv with weight `w` for x = dL/dy.
    fp32 x -> the split-operand kernel (fp32 y; `exact_w`: three weight parts = the fp32 weight exactly, the critic's
    forward convolutions); bf16 x -> the bf16 kernel (bf16 y), with hi + lo weights if `split_w`."""
    _check_act(x, "conv3x3 x")
    assert w.is_cuda and w.dtype == th.float32 and w.is_contiguous() and w.dim() == 4 and w.shape[2:] == (3, 3)
    B, C, Hin, Win = x.shape
    H, W = (2 * Hin, 2 * Win) if upsample_in else (Hin, Win)
    if dgrad:
        assert w.shape[0] == C, (w.shape, C)
        cin, cout = C, w.shape[1]
    else:
        assert w.shape[1] == C, (w.shape, C)
        cin, cout = C, w.shape[0]
    precise = x.dtype == th.float32
    flags = (FLAG_LRELU if lrelu else 0) | (FLAG_PIXELNORM if pixelnorm else 0) | \
            (FLAG_UPSAMPLE_IN if upsample_in else 0) | (FLAG_DGRAD if dgrad else 0) | \
            (FLAG_SPLIT_W if (split_w and SPLIT_W and not precise and not _forward_only[0] and H <= SPLIT_W_MAX_RES) else 0)
    y = th.empty((B, cout, H, W), dtype=x.dtype, device=x.device, memory_format=th.channels_last)
    inv = th.empty((B, H, W), dtype=th.float32, device=x.device) if (pixelnorm and want_inv_norm) else None
    l = _l()
    if bias is not None:
        assert bias.dtype == th.float32 and bias.numel() == cout and bias.is_cuda
    _account(2.0 * B * H * W * 9 * cin * cout, float(x.element_size()) * (x.numel() + y.numel()))
    if precise:
        if exact_w and not pixelnorm and H <= EXACT_W_MAX_RES:
            flags |= FLAG_W3
        kind = ("split", (1 if dgrad else 0) | (2 if flags & FLAG_W3 else 0))
        packed = _packed_weights(w, cin, cout, kind)
        with th.cuda.device(x.device):
            if packed is None:
                packed = _workspace(x.device, l.mg_conv3x3_split_workspace_bytes(cin, cout), "split_pack")
                _lib.check(l.mg_conv3x3_split_pack_weights(w.data_ptr(), cin, cout, kind[1], packed.data_ptr(), packed.numel(),
                                                           th.cuda.current_stream().cuda_stream), "mg_conv3x3_split_pack_weights")
            _lib.check(l.mg_conv3x3_split_f32(x.data_ptr(), packed.data_ptr(), bias.data_ptr() if bias is not None else None,
                                              y.data_ptr(), None, inv.data_ptr() if inv is not None else None,
                                              B, H, W, cin, cout, flags, th.cuda.current_stream().cuda_stream), "mg_conv3x3_split_f32")
        return (y, inv) if want_inv_norm else y
    mode = (1 if dgrad else 0) | (2 if flags & FLAG_SPLIT_W else 0) | (4 if pixelnorm else 0)
    packed = _packed_weights(w, cin, cout, ("bf16", mode))
    if packed is not None:
        ws, w_ptr = packed, None
    else:
        ws, w_ptr = _workspace(x.device, l.mg_conv3x3_workspace_bytes(cin, cout)), w.data_ptr()
    with th.cuda.device(x.device):
        _lib.check(l.mg_conv3x3_bf16(x.data_ptr(), w_ptr, bias.data_ptr() if bias is not None else None,
                                     y.data_ptr(), inv.data_ptr() if inv is not None else None,
                                     B, H, W, cin, cout, flags, ws.data_ptr(), ws.numel(),
                                     th.cuda.current_stream().cuda_stream), "mg_conv3x3_bf16")
    return (y, inv) if want_inv_norm else y


def conv3x3_wgrad(dy: th.Tensor, x: th.Tensor, *, upsample_in=False, out=None, accumulate=False, bias_out=None,
                  accumulate_bias=False) -> th.Tensor:
    """dw[co][ci][ky][kx] = sum_{b,y,x} dy[b,co,y,x] * xin[b,ci,y+ky-1,x+kx-1]  (fp32), xin = x or its nearest
    x2 upsampling.  `out`: write into (or, with `accumulate`, add to) an existing (Cout, Cin, 3, 3) fp32 tensor.
    `bias_out` (Cout,) fp32: also the bias gradient sum_{b,y,x} dy[b,co,y,x], from the same launch (a row of ones in the
    GEMM), written or -- `accumulate_bias` -- added."""
    # terminal product (nothing propagates from it): bf16 operands are enough, also for the fp32 layers of the precise path
    dy, x = as_act(dy), as_act(x)
    _check_act(dy, "wgrad dy"); _check_act(x, "wgrad x")
    B, cout, H, W = dy.shape
    cin = x.shape[1]
    assert x.shape[0] == B and (x.shape[2] * (2 if upsample_in else 1), x.shape[3] * (2 if upsample_in else 1)) == (H, W)
    if out is None:
        assert not accumulate
        dw = th.empty((cout, cin, 3, 3), dtype=th.float32, device=dy.device)      # overwritten: no zero fill needed
    else:
        dw = out
        assert dw.shape == (cout, cin, 3, 3) and dw.dtype == th.float32 and dw.is_contiguous() and dw.device == dy.device
    if bias_out is not None:
        assert bias_out.shape == (cout,) and bias_out.dtype == th.float32 and bias_out.is_contiguous() and bias_out.device == dy.device
    l = _l()
    _account(2.0 * B * H * W * 9 * cin * cout, 2.0 * (x.numel() + dy.numel()))
    ws = _workspace(dy.device, l.mg_conv3x3_wgrad_workspace_bytes(B, H, W, cin, cout), "wgrad")
    flags = (1 if upsample_in else 0) | (2 if accumulate else 0) | (4 if accumulate_bias else 0)
    with th.cuda.device(dy.device):
        _lib.check(l.mg_conv3x3_wgrad_bias_bf16(dy.data_ptr(), x.data_ptr(), dw.data_ptr(),
                                                bias_out.data_ptr() if bias_out is not None else None, ws.data_ptr(), ws.numel(),
                                                B, H, W, cin, cout, flags, th.cuda.current_stream().cuda_stream), "mg_conv3x3_wgrad_bias_bf16")
    return dw


# Weight gradients are LEAVES of the backward pass: nothing downstream waits for them until the optimiser step, yet
# issued in line they sit on the serial chain of a step (54 launches of a critic step, 16-25 us each on the
# low-resolution layers whatever their size, plus their reduce kernels and operand casts).  Inside a `WgradLane` the
# convolution blocks hand them to a second stream instead: the lane waits for the gradient it consumes, runs cast ->
# k_conv3x3_wgrad -> k_wgrad_reduce there, and writes (first contribution) or adds (further ones, e.g. the two
# appearances of a critic weight in the penalty's double backward) into one fp32 tensor per parameter.  The step joins
# the lane once, before the optimiser.  Under CUDA-graph capture the waits become graph edges, so the replayed graph
# runs the weight-gradient chain beside the data-gradient chain.
_lane = [None]
_lane_streams = {}
BIAS_IN_WGRAD = os.environ.get("MG_BIAS_IN_WGRAD", "1") != "0"


class WgradLane:
    def __init__(self, params, depth: int = None):
        self.params = list(params)
        self.index = {p.data_ptr(): i for i, p in enumerate(self.params) if p.requires_grad}
        self.out = {}
        self.depth = int(os.environ.get("MG_WGRAD_LANE_DEPTH", "6")) if depth is None else depth
        self.pending = collections.deque()
        self.origin = None
        self.stream = None

    def __enter__(self):
        self.origin = th.cuda.current_stream()
        key = (self.origin.device.index, self.origin.cuda_stream)
        s = _lane_streams.get(key)
        if s is None:
            s = _lane_streams[key] = th.cuda.Stream(device=self.origin.device)
        self.stream = s
        self.prev, _lane[0] = _lane[0], self
        return self

    def __exit__(self, *exc):
        _lane[0] = self.prev
        self.join()
        return False

    def accepts(self, w: th.Tensor) -> bool:
        # only first-order work (no graph is being recorded through this product) issued from the lane's own stream
        return ((not th.is_grad_enabled()) and w.dim() == 4 and tuple(w.shape[2:]) == (3, 3) and w.shape[0] >= 16 and w.shape[1] >= 16
                and w.data_ptr() in self.index and th.cuda.current_stream() == self.origin)

    def accepts_bias(self, w: th.Tensor, b) -> bool:
        """The bias gradient of the same convolution can ride on the weight-gradient launch."""
        return BIAS_IN_WGRAD and b is not None and b.dim() == 1 and b.data_ptr() in self.index and self.accepts(w)

    def _slot(self, p: th.Tensor):
        i = self.index[p.data_ptr()]
        first = i not in self.out
        if first:
            self.out[i] = th.empty(tuple(p.shape), dtype=th.float32, device=p.device)
        return self.out[i], first

    def submit(self, w: th.Tensor, g: th.Tensor, x: th.Tensor, upsample_in: bool = False, bias=None, scale=None) -> None:
        cur = self.origin
        dw, first = self._slot(w)
        db, first_b = self._slot(bias) if bias is not None else (None, True)
        self.stream.wait_stream(cur)
        with th.cuda.stream(self.stream):
            if scale is None:
                conv3x3_wgrad(g, x, upsample_in=upsample_in, out=dw, accumulate=not first, bias_out=db, accumulate_bias=not first_b)
            else:                  # the layer ran with w * scale: its contribution is scale * wgrad (two tiny kernels)
                assert bias is None
                tmp = conv3x3_wgrad(g, x, upsample_in=upsample_in)
                if first:
                    th.mul(tmp, scale, out=dw)
                else:
                    dw.addcmul_(tmp, scale)
            ev = th.cuda.Event()
            ev.record(self.stream)
        # g and x were allocated on the origin stream: they stay referenced until that stream has waited for the lane
        # (its allocator may then hand their memory to later kernels of the origin stream)
        self.pending.append((ev, g, x))
        while len(self.pending) > self.depth:
            cur.wait_event(self.pending.popleft()[0])

    def join(self) -> None:
        if self.stream is not None and (self.pending or self.out):
            self.origin.wait_stream(self.stream)
        self.pending.clear()

    def merge(self, grads):
        """`grads` of autograd.grad(..., self.params, allow_unused=True) with the lane's weight gradients filled in."""
        self.join()
        merged = []
        for i, g in enumerate(grads):
            mine = self.out.get(i)
            merged.append(mine if g is None else (g if mine is None else g + mine))
        return merged


def wgrad_lane():
    return _lane[0]


def _stream():
    return th.cuda.current_stream().cuda_stream


def _planes(x: th.Tensor, name: str) -> th.Tensor:
    if not (x.is_cuda and x.dim() == 4 and x.shape[1] == 2):
        raise TypeError(f"{name}: expected a CUDA (B, 2, H, W) tensor, got {tuple(x.shape)} on {x.device}")
    return x.float().contiguous()


def rgb_expand(x: th.Tensor, w: th.Tensor, b=None, mask_src=None, lrelu=False, out_dtype=th.bfloat16) -> th.Tensor:
    """(B,2,H,W) fp32 -> (B,C,H,W) bf16 (or fp32) channels_last: W x + b, then LeakyReLU(0.2) or a mask multiply."""
    x = _planes(x, "rgb_expand x")
    B, _, H, W = x.shape
    C = w.shape[0]
    w = w.float().reshape(C, 2).contiguous()
    if mask_src is not None:
        out_dtype = mask_src.dtype
    y = th.empty((B, C, H, W), dtype=out_dtype, device=x.device, memory_format=th.channels_last)
    mode = 1 if lrelu else (2 if mask_src is not None else 0)
    if mask_src is not None:
        _check_act(mask_src, "rgb_expand mask")
    with th.cuda.device(x.device):
        _lib.check(getattr(_l(), "mg_rgb_expand" + _sfx(y))(x.data_ptr(), w.data_ptr(), b.float().contiguous().data_ptr() if b is not None else None,
                                           mask_src.data_ptr() if mask_src is not None else None, y.data_ptr(),
                                           B, H * W, C, mode, _stream()), "mg_rgb_expand_bf16")
    return y


def rgb_project(a: th.Tensor, w: th.Tensor, bias=None, mask_src=None, tanh=False, w_is_c_by_2=False) -> th.Tensor:
    """(B,C,H,W) bf16 channels_last -> (B,2,H,W) fp32.  `w` is (2,C) (forward ToMagnPhase weight) or, with
    w_is_c_by_2, a (C,2) weight used transposed (data gradient of rgb_expand)."""
    _check_act(a, "rgb_project a")
    B, C, H, W = a.shape
    w = w.float().reshape(C, 2).contiguous() if w_is_c_by_2 else w.float().reshape(2, C).contiguous()
    rs, cs = (1, 2) if w_is_c_by_2 else (C, 1)
    out = th.empty((B, 2, H, W), dtype=th.float32, device=a.device)
    if mask_src is not None:
        _check_act(mask_src, "rgb_project mask", a.dtype)
    with th.cuda.device(a.device):
        _lib.check(getattr(_l(), "mg_rgb_project" + _sfx(a))(a.data_ptr(), w.data_ptr(), rs, cs, bias.float().contiguous().data_ptr() if bias is not None else None,
                                            mask_src.data_ptr() if mask_src is not None else None, out.data_ptr(),
                                            B, H * W, C, 1 if tanh else 0, _stream()), "mg_rgb_project_bf16")
    return out


def rgb_wgrad(g: th.Tensor, mask_src, x: th.Tensor):
    """(gw (C,2) fp32, gb (C,) fp32) = sums over pixels of (g * mask) (x) x and of (g * mask)."""
    _check_act(g, "rgb_wgrad g")
    x = _planes(x, "rgb_wgrad x")
    B, C, H, W = g.shape
    gw = th.zeros((C, 2), dtype=th.float32, device=g.device)
    gb = th.zeros((C,), dtype=th.float32, device=g.device)
    if mask_src is not None:
        _check_act(mask_src, "rgb_wgrad mask", g.dtype)
    with th.cuda.device(g.device):
        _lib.check(getattr(_l(), "mg_rgb_wgrad" + _sfx(g))(g.data_ptr(), mask_src.data_ptr() if mask_src is not None else None, x.data_ptr(),
                                          gw.data_ptr(), gb.data_ptr(), B, H * W, C, _stream()), "mg_rgb_wgrad_bf16")
    return gw, gb


def pool2(x: th.Tensor, adjoint: bool = False, sum_pool: bool = False) -> th.Tensor:
    """AvgPool2d(2,2) on bf16 channels_last, its adjoint (0.25 * nearest x2 replication), or the 2x2 sum pooling
    (backward of a nearest x2 upsampling)."""
    _check_act(x, "pool2 x")
    B, C, H, W = x.shape
    if adjoint:
        ho, wo = H, W
        out = th.empty((B, C, 2 * H, 2 * W), dtype=x.dtype, device=x.device, memory_format=th.channels_last)
    else:
        assert H % 2 == 0 and W % 2 == 0
        ho, wo = H // 2, W // 2
        out = th.empty((B, C, ho, wo), dtype=x.dtype, device=x.device, memory_format=th.channels_last)
    with th.cuda.device(x.device):
        _lib.check(getattr(_l(), "mg_pool2" + _sfx(x))(x.data_ptr(), out.data_ptr(), B, ho, wo, C, 1 if adjoint else (2 if sum_pool else 0), _stream()),
                   "mg_pool2_bf16")
    return out


def pixelnorm_lrelu_bwd(go: th.Tensor, o: th.Tensor, inv: th.Tensor, want_bias_grad: bool = True):
    """Backward of LeakyReLU -> PixelNorm (saved normalised output `o`, saved 1/norm `inv`): (gz bf16, gb fp32)."""
    _check_act(o, "pixelnorm_lrelu_bwd o")
    go = as_act(go, o.dtype)
    B, C, H, W = o.shape
    inv = inv.float().contiguous()
    gz = th.empty_like(o)
    gb = th.empty((C,), dtype=th.float32, device=o.device) if want_bias_grad else None      # overwritten by the kernel
    l = _l()
    ws = _workspace(o.device, l.mg_colsum_workspace_bytes(C), "colsum") if want_bias_grad else None
    with th.cuda.device(o.device):
        _lib.check(getattr(l, "mg_pixelnorm_lrelu_bwd" + _sfx(o))(go.data_ptr(), o.data_ptr(), inv.data_ptr(), gz.data_ptr(),
                                                 gb.data_ptr() if gb is not None else None,
                                                 ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0,
                                                 B * H * W, C, _stream()),
                   "mg_pixelnorm_lrelu_bwd_bf16")
    return gz, gb


def lrelu_bwd(gy: th.Tensor, y: th.Tensor, want_bias_grad: bool = True):
    """gz = gy * (y > 0 ? 1 : 0.2) and, fused, the bias gradient sum over pixels (fp32)."""
    _check_act(y, "lrelu_bwd y")
    gy = as_act(gy, y.dtype)
    B, C, H, W = y.shape
    gz = th.empty_like(y)
    gb = th.empty((C,), dtype=th.float32, device=y.device) if want_bias_grad else None      # overwritten by the kernel
    l = _l()
    ws = _workspace(y.device, l.mg_colsum_workspace_bytes(C), "colsum") if want_bias_grad else None
    with th.cuda.device(y.device):
        _lib.check(getattr(l, "mg_lrelu_bwd" + _sfx(y))(gy.data_ptr(), y.data_ptr(), gz.data_ptr(), gb.data_ptr() if gb is not None else None,
                                       ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0,
                                       B * H * W, C, _stream()), "mg_lrelu_bwd_bf16")
    return gz, gb


def unpool_lrelu_bwd(gp: th.Tensor, h: th.Tensor, want_bias_grad: bool = True):
    """Backward of LeakyReLU -> AvgPool2d(2,2) in one pass: gz = 0.25 * up2(gp) * (h > 0 ? 1 : 0.2) at the resolution of
    `h`, and the bias gradient sum over pixels (fp32)."""
    _check_act(h, "unpool_lrelu_bwd h")
    gp = as_act(gp, h.dtype)
    B, C, H, W = h.shape
    assert gp.shape == (B, C, H // 2, W // 2) and H % 2 == 0 and W % 2 == 0, (gp.shape, h.shape)
    gz = th.empty_like(h)
    gb = th.empty((C,), dtype=th.float32, device=h.device) if want_bias_grad else None
    l = _l()
    ws = _workspace(h.device, l.mg_colsum_workspace_bytes(C), "colsum") if want_bias_grad else None
    with th.cuda.device(h.device):
        _lib.check(getattr(l, "mg_unpool2_lrelu_bwd" + _sfx(h))(gp.data_ptr(), h.data_ptr(), gz.data_ptr(), gb.data_ptr() if gb is not None else None,
                                               ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0,
                                               B, H // 2, W // 2, C, _stream()), "mg_unpool2_lrelu_bwd_bf16")
    return gz, gb


def pool2_planes(x: th.Tensor, adjoint: bool = False) -> th.Tensor:
    """AvgPool2d(2, 2) of fp32 NCHW planes (B, C, H, W) -> (B, C, H/2, W/2), or (adjoint) its backward
    (B, C, H, W) -> (B, C, 2H, 2W) = 0.25 * x replicated.  Same bits as torch's avg_pool2d / avg_pool2d_backward."""
    x = _planes(x, "pool2_planes x")
    B, C, H, W = x.shape
    if adjoint:
        out = th.empty((B, C, 2 * H, 2 * W), dtype=th.float32, device=x.device)
        ho, wo = H, W
    else:
        assert H % 2 == 0 and W % 2 == 0
        ho, wo = H // 2, W // 2
        out = th.empty((B, C, ho, wo), dtype=th.float32, device=x.device)
    with th.cuda.device(x.device):
        _lib.check(_l().mg_pool2_planes_f32(x.data_ptr(), out.data_ptr(), B * C, ho, wo, 1 if adjoint else 0, _stream()),
                   "mg_pool2_planes_f32")
    return out
