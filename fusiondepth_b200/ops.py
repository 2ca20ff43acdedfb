"""torch.autograd.Function wrappers over the C-ABI kernels.

PyTorch here is plumbing only: it owns device memory (torch.empty), the stream and the autograd
tape.  Activations are fp32 tensors of logical shape [B,C,H,W] stored channels-last (NHWC); conv
weights are logical [Cout,Cin,KH,KW] stored channels-last ([Cout,KH,KW,Cin]).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_void_p
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib

ACT = {"none": 0, "relu": 1, "elu": 2, "sigmoid": 3, "tanh": 4}
# "tc": tcgen05 tensor-core convolutions wherever eligible (default); "cudacore": exact-fp32 path
CONV_BACKEND = os.environ.get("FD_CONV", "tc")

# Set by training.TrainStep: parameters carry a preallocated gradient view (`p._fd_grad`, a slice of
# the flat gradient buffer) and the backward kernels add into it directly (the weight-gradient and
# bias kernels accumulate with reductions anyway), so autograd's AccumulateGrad pass and the zeroed
# temporaries disappear.
DIRECT_GRAD = False

# Per-step cache of derived weight tensors (W_lo, transposed W / W_lo, padded stem weights), keyed by
# the weight's address.  TrainStep clears it at the start of every optimiser step; entries carry the
# event of the stream that produced them so other trunk streams can wait on it.
WEIGHT_CACHE: Optional[Dict] = None


class StatsArena:
    """Zeroed fp64 scratch for the conv-epilogue BatchNorm statistics of one optimiser step.  Every conv + BN pair
    used to allocate and zero its own [2*C] buffer (228 FillFunctor launches per ResNet-18 step); TrainStep
    zeroes this arena ONCE at the start of the step and the pairs take consecutive 256-byte-aligned slices (host
    side bump pointer: the call order is the same every step, so a captured graph sees fixed addresses)."""

    def __init__(self, device, nbytes: int = 16 << 20):
        self.buf = torch.zeros(nbytes // 8, device=device, dtype=torch.float64)
        self.off = 0

    def begin_step(self):
        self.buf.zero_()
        self.off = 0

    def take(self, n: int):
        n_al = (n + 31) // 32 * 32
        if self.off + n_al > self.buf.numel():
            return None
        t = self.buf[self.off:self.off + n]
        self.off += n_al
        return t


STATS_ARENA: Optional[StatsArena] = None


def zeroed_stats(n: int, device) -> torch.Tensor:
    """A zeroed float64 [n] buffer for fd_conv2d_fwd_tc_stats: a slice of the step's arena when one is active."""
    # Opt-in (FD_STATS_ARENA=1): measured no faster than the per-call fills (504 vs 503-505 images/s).  (The 5 %
    # graph-replay-vs-eager gradient mismatch once seen with it was the skip-connection gradient race described in
    # AssembleFn.backward, which any change of timing could expose -- not the arena.)
    if STATS_ARENA is not None and os.environ.get("FD_STATS_ARENA", "0") == "1":
        t = STATS_ARENA.take(n)
        if t is not None:
            return t
    return torch.zeros(n, device=device, dtype=torch.float64)


def _direct_grad(p):
    return getattr(p, "_fd_grad", None) if (DIRECT_GRAD and p is not None) else None


def _cached(key, make, persistent=True):
    """make() -> tuple of tensors, computed once per step when the cache is on.  The key is a device
    address, so only tensors that live for the whole step (parameters) may insert an entry: a temporary's
    address can be handed to another layer by the allocator within the same step (`persistent=False`:
    look up, never insert)."""
    if WEIGHT_CACHE is None:
        return make()
    hit = WEIGHT_CACHE.get(key)
    cur = torch.cuda.current_stream()
    if hit is None:
        val = make()
        if persistent:
            ev = torch.cuda.Event()
            ev.record(cur)
            WEIGHT_CACHE[key] = (val, ev, cur)
        return val
    val, ev, owner = hit
    if owner != cur:
        cur.wait_event(ev)
    return val
CL = torch.channels_last


class BNSchedule:
    """Order-free BatchNorm running-statistic updates for one optimiser step.

    nn.BatchNorm2d applies r <- (1-m) r + m s once per call, in call order.  TrainStep runs the
    micro-batches (and the two pose-pair calls of a trunk) concurrently on different streams, so it
    uses the closed form instead: n calls give r <- (1-m)^n r + sum_k m (1-m)^(n-1-k) s_k.
    begin_step() decays every buffer once (and bumps num_batches_tracked by n); call k of a layer
    then adds its weighted statistic atomically (fd_bn_fwd stat_weight >= 0)."""

    def __init__(self, calls_per_step: Dict):
        self.n = dict(calls_per_step)               # BatchNorm2d module -> calls per optimiser step
        self.count: Dict = {}
        groups: Dict = {}
        for mod, n in self.n.items():
            groups.setdefault((float(mod.momentum), int(n)), []).append(mod)
        self.groups = [((1.0 - m) ** n, n, [b for mod in mods for b in (mod.running_mean, mod.running_var)],
                        [mod.num_batches_tracked for mod in mods]) for (m, n), mods in groups.items()]

    def begin_step(self):
        self.count = {}
        for decay, n, bufs, nbt in self.groups:
            torch._foreach_mul_(bufs, decay)
            torch._foreach_add_(nbt, n)

    def next_weight(self, mod) -> float:
        n = self.n.get(mod)
        if n is None:
            return -1.0
        k = self.count.get(mod, 0)
        if k >= n:
            raise RuntimeError("BNSchedule: BatchNorm layer called %d times in one step, scheduled for %d" % (k + 1, n))
        self.count[mod] = k + 1
        m = float(mod.momentum)
        return m * (1.0 - m) ** (n - 1 - k)

    def end_step(self):
        bad = [(self.count.get(mod, 0), n) for mod, n in self.n.items() if self.count.get(mod, 0) != n]
        if bad:
            raise RuntimeError("BNSchedule: %d BatchNorm layers were not called as scheduled, e.g. %s" % (len(bad), bad[0]))


# Set by training.TrainStep while a step is being issued; None => nn.BatchNorm2d's in-place update.
BN_SCHEDULE: Optional[BNSchedule] = None


# Weight-gradient kernels off the critical path.  Set by training.TrainStep to a dict {stream -> side stream}:
# Conv2dFn.backward then launches the weight-gradient kernel on the side stream of the stream it runs on (forked
# by an event once dy is final), so the data-gradient chain of the backward pass -- the critical path of the
# step -- no longer waits for it; nothing consumes dW before the gradient exchange.  Only with DIRECT_GRAD (the
# kernel accumulates into the flat gradient buffer, autograd never sees dW).  WGRAD_KEEP holds x / dy until the
# step joins the side streams, so the caching allocator cannot hand their memory to the main stream early.
WGRAD_STREAMS: Optional[Dict] = None
WGRAD_USED: set = set()
WGRAD_KEEP: list = []


def wgrad_side_stream():
    if WGRAD_STREAMS is None:
        return None
    cur = torch.cuda.current_stream()
    side = WGRAD_STREAMS.get(cur)
    if side is None:
        side = WGRAD_STREAMS[cur] = torch.cuda.Stream(device=cur.device)
    return side


def join_wgrad_streams():
    """The current stream waits for every weight-gradient side stream used since the last join."""
    cur = torch.cuda.current_stream()
    for s in WGRAD_USED:
        cur.wait_stream(s)
    WGRAD_USED.clear()
    WGRAD_KEEP.clear()


# Set by training.TrainStep when the gradient exchange is bucketed: callable(bucket_id), invoked from the
# backward pass of every trunk once the gradients of that bucket have been launched (grad_ready()).
GRAD_READY = None


# Optional per-kernel-family timing (bench.py's roofline leg): when PROFILE is a dict, every
# timed call appends (start_event, end_event, algorithmic_work) under its family name.  Events are
# recorded on the launching stream; nothing is synchronised here.
PROFILE: Optional[Dict] = None


class _timed:
    def __init__(self, name: str, work: float):
        self.name, self.work = name, work

    def __enter__(self):
        if PROFILE is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            PROFILE.setdefault(self.name, []).append((self.s, e, self.work))
        return False


def _p(t: Optional[torch.Tensor]):
    return None if t is None else c_void_p(t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t: torch.Tensor, who: str):
    if not t.is_cuda:
        raise _lib.FusionDepthLibraryError(
            "%s: fusiondepth_b200 operators run on CUDA only (got a %s tensor); there is no CPU "
            "fallback" % (who, t.device))
    if t.dtype != torch.float32:
        raise TypeError("%s: expected float32, got %s" % (who, t.dtype))


def nhwc(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous(memory_format=CL)


def empty_nhwc(B, C, H, W, device):
    return torch.empty((B, C, H, W), device=device, dtype=torch.float32, memory_format=CL)


# --------------------------------------------------------------------------------------------
class PrepInput(torch.autograd.Function):
    """(x - 0.45)/0.225 and NCHW -> NHWC (resnet_encoder.py:94).  Inputs carry no gradient."""

    @staticmethod
    def forward(ctx, x):
        _require_cuda(x, "prep_input")
        x = x.contiguous()
        B, C, H, W = x.shape
        y = empty_nhwc(B, C, H, W, x.device)
        _lib.check(_lib.load().fd_prep_input(_p(x), _p(y), B, C, H, W, 0.45, 0.225, _stream()),
                   "fd_prep_input")
        ctx.mark_non_differentiable(y)
        return y

    @staticmethod
    def backward(ctx, g):
        return None


def prep_input(x):
    return PrepInput.apply(x)


# --------------------------------------------------------------------------------------------
# BatchNorm batch statistics out of the tensor-core convolution's epilogue instead of a separate pass over the
# conv output (FD_BN_FUSE_STATS=0 restores the separate bn_stats kernel; both are parity-tested).
FUSE_BN_STATS = os.environ.get("FD_BN_FUSE_STATS", "1") == "1"
# BatchNorm + ReLU backward without the saved output (mask re-evaluated from x); FD_BN_XMASK=0 reads y back
BN_XMASK = os.environ.get("FD_BN_XMASK", "1") == "1"


def conv_emits_stats(Cin, Cout, has_bias) -> bool:
    """True when conv2d(..., stats=ws) will fill `ws` (tensor-core conv_tc2 forward path)."""
    return (FUSE_BN_STATS and CONV_BACKEND == "tc" and not has_bias and Cin % 32 == 0 and Cout % 16 == 0
            and os.environ.get("FD_CONV_TC", "") != "v1" and not (Cout == 16 and Cin == 16))


class Conv2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, act, stats=None):
        _require_cuda(x, "conv2d")
        lib = _lib.load()
        x = nhwc(x)
        w = nhwc(weight)
        # derived tensors may be cached by address only for parameters stored channels-last in place
        ctx.param_w = isinstance(weight, torch.nn.Parameter) and w.data_ptr() == weight.data_ptr()
        B, Cin, H, W = x.shape
        Cout, Cin_w, KH, KW = w.shape
        if Cin_w != Cin:
            raise RuntimeError("conv2d: input has %d channels, weight expects %d" % (Cin, Cin_w))
        Ho = (H + 2 * pad - KH) // stride + 1
        Wo = (W + 2 * pad - KW) // stride + 1
        y = empty_nhwc(B, Cout, Ho, Wo, x.device)
        use_tc = CONV_BACKEND == "tc" and Cin % 32 == 0 and Cout % 16 == 0
        ctx.cout1 = bool(lib.fd_conv2d_cout1_supported(Cin, Cout, KH, KW, stride)) and CONV_BACKEND != "cudacore"
        ctx.c16 = (bool(lib.fd_conv2d_c16_supported(Cin, Cout, KH, KW, stride)) and CONV_BACKEND == "tc"
                   and os.environ.get("FD_CONV16", "1") != "0")
        if ctx.c16 and Cin == 16:
            # 16 -> 16 channels at full resolution: direct CUDA-core convolution (conv_small.cu)
            with _timed("conv", 2.0 * B * Ho * Wo * Cout * KH * KW * Cin):
                _lib.check(lib.fd_conv2d_c16_fwd(_p(x), _p(w), _p(bias), _p(y), B, H, W, Cin, pad, act,
                                                 _stream()), "fd_conv2d_c16_fwd")
        elif ctx.cout1:
            # disparity heads: one output channel, streaming kernels (conv_small.cu)
            with _timed("conv", 2.0 * B * Ho * Wo * Cout * KH * KW * Cin):
                _lib.check(lib.fd_conv2d_cout1_fwd(_p(x), _p(w), _p(bias), _p(y), B, H, W, Cin, pad, act,
                                                   _stream()), "fd_conv2d_cout1_fwd")
        elif use_tc:
            def make_lo():
                t = torch.empty(w.numel(), device=x.device, dtype=torch.float32)
                _lib.check(lib.fd_tf32_split(_p(w), _p(t), w.numel(), _stream()), "fd_tf32_split")
                return (t,)
            (wlo,) = _cached((w.data_ptr(), "lo"), make_lo, ctx.param_w)
            with _timed("conv", 2.0 * B * Ho * Wo * Cout * KH * KW * Cin):
                _lib.check(lib.fd_conv2d_fwd_tc_stats(_p(x), _p(w), _p(wlo), _p(bias), _p(y), B, H, W, Cin,
                                                      Cout, KH, KW, stride, pad, act, _p(stats), _stream()),
                           "fd_conv2d_fwd_tc")
            stats = None
        else:
            with _timed("conv", 2.0 * B * Ho * Wo * Cout * KH * KW * Cin):
                _lib.check(lib.fd_conv2d_fwd(_p(x), _p(w), _p(bias), _p(y), B, H, W, Cin, Cout, KH, KW,
                                             stride, pad, act, _stream()), "fd_conv2d_fwd")
        if stats is not None:
            raise RuntimeError("conv2d: channel statistics were requested on a path that does not produce them")
        ctx.save_for_backward(x, w, y if act != 0 else None)
        ctx.cfg = (stride, pad, act, bias is not None)
        ctx.wg, ctx.bg = _direct_grad(weight), _direct_grad(bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, w, y = ctx.saved_tensors
        stride, pad, act, has_bias = ctx.cfg
        B, Cin, H, W = x.shape
        Cout, _, KH, KW = w.shape
        dy = nhwc(dy)
        M = dy.shape[0] * dy.shape[2] * dy.shape[3]
        st = _stream()
        dbias = None
        if act != 0 or has_bias:
            if has_bias:
                dbias = ctx.bg if ctx.bg is not None else torch.zeros(Cout, device=x.device, dtype=torch.float32)
            dpre = torch.empty_like(dy) if act != 0 else dy
            _lib.check(lib.fd_act_bwd(_p(y if act != 0 else dy), _p(dy), _p(dpre), _p(dbias), M, Cout,
                                      act, st), "fd_act_bwd")
            dy = dpre
        dx = None
        if ctx.cout1 or ctx.c16:
            dgrad_fn, wgrad_fn, who = ((lib.fd_conv2d_cout1_dgrad, lib.fd_conv2d_cout1_wgrad, "fd_conv2d_cout1")
                                       if ctx.cout1 else
                                       (lib.fd_conv2d_c16_dgrad, lib.fd_conv2d_c16_wgrad, "fd_conv2d_c16"))
            if ctx.needs_input_grad[0]:
                dx = empty_nhwc(B, Cin, H, W, x.device)
                with _timed("conv", 2.0 * M * Cout * KH * KW * Cin):
                    _lib.check(dgrad_fn(_p(dy), _p(w), _p(dx), B, H, W, Cin, pad, st), who + "_dgrad")
            dw = None
            if ctx.needs_input_grad[1]:
                dw = ctx.wg if ctx.wg is not None else torch.empty(
                    (Cout, Cin, KH, KW), device=x.device, dtype=torch.float32, memory_format=CL).zero_()
                with _timed("conv", 2.0 * M * Cout * KH * KW * Cin):
                    _lib.check(wgrad_fn(_p(x), _p(dy), _p(dw), B, H, W, Cin, pad, st), who + "_wgrad")
            return (dx, None if ctx.wg is not None else dw, None if ctx.bg is not None else dbias,
                    None, None, None, None)
        if ctx.needs_input_grad[0]:
            dx = empty_nhwc(B, Cin, H, W, x.device)
            if CONV_BACKEND == "tc" and Cout % 32 == 0 and Cin % 16 == 0:
                def make_t():
                    a = torch.empty(w.numel(), device=x.device, dtype=torch.float32)
                    b = torch.empty(w.numel(), device=x.device, dtype=torch.float32)
                    _lib.check(lib.fd_weight_transpose_split(_p(w), _p(a), _p(b), Cout, KH * KW, Cin,
                                                             _stream()), "fd_weight_transpose_split")
                    return a, b
                wt, wtlo = _cached((w.data_ptr(), "t"), make_t, ctx.param_w)
                with _timed("conv", 2.0 * M * Cout * KH * KW * Cin):
                    _lib.check(lib.fd_conv2d_dgrad_tc(_p(dy), _p(wt), _p(wtlo), _p(dx), B, H, W, Cin,
                                                      Cout, KH, KW, stride, pad, st), "fd_conv2d_dgrad_tc")
            else:
                wt = torch.empty(w.numel(), device=x.device, dtype=torch.float32)
                _lib.check(lib.fd_weight_transpose(_p(w), _p(wt), Cout, KH * KW, Cin, st),
                           "fd_weight_transpose")
                with _timed("conv", 2.0 * M * Cout * KH * KW * Cin):
                    _lib.check(lib.fd_conv2d_dgrad(_p(dy), _p(wt), _p(dx), B, H, W, Cin, Cout, KH, KW,
                                                   stride, pad, st), "fd_conv2d_dgrad")
        dw = None
        if ctx.needs_input_grad[1]:
            dw = ctx.wg if ctx.wg is not None else torch.empty(
                (Cout, Cin, KH, KW), device=x.device, dtype=torch.float32, memory_format=CL).zero_()
            side = wgrad_side_stream() if (ctx.wg is not None and PROFILE is None) else None
            wst = st
            if side is not None:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())       # dy is final; (the data gradient above only reads it)
                side.wait_event(ev)
                WGRAD_USED.add(side)
                WGRAD_KEEP.append((x, dy))
                wst = c_void_p(side.cuda_stream)
            with _timed("conv", 2.0 * M * Cout * KH * KW * Cin):
                if CONV_BACKEND == "tc" and Cin % 32 == 0 and Cout % 32 == 0:
                    _lib.check(lib.fd_conv2d_wgrad_tc(_p(x), _p(dy), _p(dw), B, H, W, Cin, Cout, KH, KW,
                                                      stride, pad, wst), "fd_conv2d_wgrad_tc")
                else:
                    _lib.check(lib.fd_conv2d_wgrad(_p(x), _p(dy), _p(dw), B, H, W, Cin, Cout, KH, KW,
                                                   stride, pad, wst), "fd_conv2d_wgrad")
        # gradients written straight into the parameters' buffers are not handed back to autograd
        return (dx, None if ctx.wg is not None else dw, None if ctx.bg is not None else dbias,
                None, None, None, None)


class PadConvParamsFn(torch.autograd.Function):
    """Zero-pads a conv weight [Cout,Cin,KH,KW] (and bias) to [Cout_p,Cin_p,KH,KW] so that layers with channel
    counts that are not multiples of 32 (the refine2d decoder's 262 / 134 / 102 / 22) run
    on the tensor-core kernels; the gradient of the padded copy is folded back into the parameter's."""

    @staticmethod
    def forward(ctx, weight, bias, cout_p, cin_p):
        lib = _lib.load()
        w = nhwc(weight)
        Cout, Cin, KH, KW = w.shape
        st = _stream()
        wp = torch.empty((cout_p, cin_p, KH, KW), device=w.device, dtype=torch.float32, memory_format=CL)
        if cout_p > Cout:
            wp.zero_()
        _lib.check(lib.fd_pad_rows(_p(w), _p(wp), Cout * KH * KW, Cin, cin_p, 0, st), "fd_pad_rows")
        bp = None
        if bias is not None:
            bp = torch.zeros(cout_p, device=w.device, dtype=torch.float32)
            bp[:Cout].copy_(bias)
        ctx.cfg = (Cout, Cin, KH, KW, cout_p, cin_p, bias is not None)
        ctx.wg, ctx.bg = _direct_grad(weight), _direct_grad(bias)
        return wp, bp

    @staticmethod
    def backward(ctx, dwp, dbp):
        lib = _lib.load()
        Cout, Cin, KH, KW, cout_p, cin_p, has_bias = ctx.cfg
        dwp = nhwc(dwp)
        st = _stream()
        if ctx.wg is not None:
            _lib.check(lib.fd_pad_rows(_p(dwp), _p(ctx.wg), Cout * KH * KW, cin_p, Cin, 1, st), "fd_pad_rows")
            dw = None
        else:
            dw = torch.empty((Cout, Cin, KH, KW), device=dwp.device, dtype=torch.float32, memory_format=CL).zero_()
            _lib.check(lib.fd_pad_rows(_p(dwp), _p(dw), Cout * KH * KW, cin_p, Cin, 1, st), "fd_pad_rows")
        db = None
        if has_bias and dbp is not None:
            if ctx.bg is not None:
                ctx.bg.add_(dbp[:Cout])
            else:
                db = dbp[:Cout].clone()
        return dw, db, None, None


def pad_conv_params(weight, bias, cout_p, cin_p):
    return PadConvParamsFn.apply(weight, bias, int(cout_p), int(cin_p))


_ZERO_SEGMENTS: Dict = {}


def zero_segment(B, C, H, W, device):
    """A constant all-zero [B,C,H,W] channels-last tensor (channel padding for assemble())."""
    key = (B, C, H, W, str(device))
    z = _ZERO_SEGMENTS.get(key)
    if z is None:
        z = _ZERO_SEGMENTS[key] = torch.zeros((B, C, H, W), device=device, dtype=torch.float32).contiguous(
            memory_format=CL)
    return z


class GradReadyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bucket):
        ctx.bucket = bucket
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if GRAD_READY is not None:
            GRAD_READY(ctx.bucket)
        return g, None


def grad_ready(x, bucket: int):
    """Marks the point of the forward pass after which all parameters belong to `bucket`'s predecessors in
    backward order; a no-op unless a bucketed gradient exchange is active."""
    if GRAD_READY is None or not (torch.is_grad_enabled() and x.requires_grad):
        return x
    return GradReadyFn.apply(x, int(bucket))


def conv2d(x, weight, bias=None, stride=1, pad=0, act="none", stats=None):
    """stats: optional zeroed float64 [2*Cout] tensor that receives the per-channel sum / sum of squares
    of the output (only on the path conv_emits_stats() describes)."""
    return Conv2dFn.apply(x, weight, bias, int(stride), int(pad), ACT[act], stats)


# --------------------------------------------------------------------------------------------
class StemConvFn(torch.autograd.Function):
    """conv1 of the ResNet trunks (7x7/2, pad 3, Cin 2..6, no bias) with the input normalisation
    folded in: normalised im2col rows -> tensor-core GEMM (resnet_encoder.py:94-95).  The image
    carries no gradient; the weight gradient is a tensor-core GEMM over the saved rows."""

    @staticmethod
    def forward(ctx, x, weight, bias=None, act=0):
        _require_cuda(x, "stem_conv")
        lib = _lib.load()
        x = x.contiguous()
        w = nhwc(weight)
        B, C, H, W = x.shape
        Cout, Cw, KH, KW = w.shape
        if Cw != C:
            raise RuntimeError("stem_conv: input has %d channels, weight expects %d" % (C, Cw))
        stride, pad = 2, 3
        Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
        K = KH * KW * C
        Kpad = (K + 31) // 32 * 32
        st = _stream()
        A = torch.empty((B * Ho * Wo, Kpad), device=x.device, dtype=torch.float32)
        _lib.check(lib.fd_stem_im2col(_p(x), _p(A), B, C, H, W, KH, KW, stride, pad, Kpad, 0.45, 0.225, st),
                   "fd_stem_im2col")
        def make_pad():
            a = torch.empty((Cout, Kpad), device=x.device, dtype=torch.float32)
            _lib.check(lib.fd_pad_rows(_p(w), _p(a), Cout, K, Kpad, 0, _stream()), "fd_pad_rows")
            b = torch.empty_like(a)
            _lib.check(lib.fd_tf32_split(_p(a), _p(b), a.numel(), _stream()), "fd_tf32_split")
            return a, b
        wpad, wlo = _cached((w.data_ptr(), "stem"), make_pad)
        y = empty_nhwc(B, Cout, Ho, Wo, x.device)
        with _timed("conv", 2.0 * B * Ho * Wo * Cout * K):
            _lib.check(lib.fd_conv2d_fwd_tc(_p(A), _p(wpad), _p(wlo), _p(bias), _p(y), B, Ho, Wo, Kpad, Cout,
                                            1, 1, 1, 0, act, st), "fd_conv2d_fwd_tc")
        ctx.fused_epilogue = bias is not None or act != 0
        ctx.save_for_backward(A)
        ctx.cfg = (B, Ho, Wo, Kpad, Cout, C, KH, KW, K)
        ctx.wg = _direct_grad(weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        if ctx.fused_epilogue:
            raise NotImplementedError("stem_conv: bias / activation are inference-only (folded BatchNorm)")
        (A,) = ctx.saved_tensors
        B, Ho, Wo, Kpad, Cout, C, KH, KW, K = ctx.cfg
        dy = nhwc(dy)
        st = _stream()
        dwpad = torch.zeros((Cout, Kpad), device=dy.device, dtype=torch.float32)
        with _timed("conv", 2.0 * B * Ho * Wo * Cout * K):
            _lib.check(lib.fd_conv2d_wgrad_tc(_p(A), _p(dy), _p(dwpad), B, Ho, Wo, Kpad, Cout, 1, 1, 1, 0, st),
                       "fd_conv2d_wgrad_tc")
        if ctx.wg is not None:
            _lib.check(lib.fd_pad_rows(_p(dwpad), _p(ctx.wg), Cout, Kpad, K, 1, st), "fd_pad_rows")
            return None, None, None, None
        dw = torch.empty((Cout, C, KH, KW), device=dy.device, dtype=torch.float32, memory_format=CL)
        _lib.check(lib.fd_pad_rows(_p(dwpad), _p(dw), Cout, Kpad, K, 0, st), "fd_pad_rows")
        return None, dw, None, None


def stem_conv(x, weight, bias=None, act="none"):
    """(x-0.45)/0.225 -> 7x7/2 conv, NHWC output.  Tensor-core path when enabled.  bias / act: inference only
    (a BatchNorm folded into the stem)."""
    Cout, C, KH, KW = weight.shape
    if CONV_BACKEND == "tc" and (KH, KW) == (7, 7) and Cout % 32 == 0:
        return StemConvFn.apply(x, weight, bias, ACT[act])
    return conv2d(prep_input(x), weight, bias, 2, 3, act)


# --------------------------------------------------------------------------------------------
class BatchNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, residual, training, momentum, eps,
                relu, stat_weight, stats=None):
        _require_cuda(x, "batch_norm")
        lib = _lib.load()
        x = nhwc(x)
        if residual is not None:
            residual = nhwc(residual)
        B, C, H, W = x.shape
        M = B * H * W
        y = torch.empty_like(x)
        mean = torch.empty(C, device=x.device, dtype=torch.float32)
        rstd = torch.empty(C, device=x.device, dtype=torch.float32)
        ws = stats if stats is not None else torch.empty(lib.fd_bn_workspace_bytes(C) // 8, device=x.device,
                                                         dtype=torch.float64)
        _lib.check(lib.fd_bn_fwd(_p(x), _p(residual), _p(gamma), _p(beta), _p(running_mean),
                                 _p(running_var), int(training), momentum, eps, int(relu), _p(y),
                                 _p(mean), _p(rstd), _p(ws), M, C, stat_weight, int(stats is not None),
                                 _stream()), "fd_bn_fwd")
        # residual-free BatchNorm + ReLU in training: the backward rebuilds the ReLU mask from x (fd_bn_bwd_xmask)
        # instead of reading y back -- beta is saved in its place
        xmask = bool(BN_XMASK and relu and training and residual is None and lib.fd_bn_bwd_xmask_ok(M, C))
        ctx.save_for_backward(x, beta if xmask else y, gamma, mean, rstd)
        ctx.cfg = (int(relu), int(training), residual is not None, xmask)
        ctx.gg, ctx.gb = _direct_grad(gamma), _direct_grad(beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, y, gamma, mean, rstd = ctx.saved_tensors
        relu, training, has_res, xmask = ctx.cfg
        dy = nhwc(dy)
        B, C, H, W = x.shape
        M = B * H * W
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if has_res else None
        direct = ctx.gg is not None and ctx.gb is not None
        dgamma = ctx.gg if direct else torch.empty(C, device=x.device, dtype=torch.float32)
        dbeta = ctx.gb if direct else torch.empty(C, device=x.device, dtype=torch.float32)
        ws = torch.empty(lib.fd_bn_workspace_bytes(C) // 8, device=x.device, dtype=torch.float64)
        if xmask:       # `y` holds beta
            _lib.check(lib.fd_bn_bwd_xmask(_p(x), _p(dy), _p(gamma), _p(y), _p(mean), _p(rstd), training, _p(dx),
                                           _p(dgamma), _p(dbeta), _p(ws), M, C, int(direct), _stream()),
                       "fd_bn_bwd_xmask")
        else:
            _lib.check(lib.fd_bn_bwd(_p(x), _p(y), _p(dy), _p(gamma), _p(mean), _p(rstd), relu, training,
                                     _p(dx), _p(dres), _p(dgamma), _p(dbeta), _p(ws), M, C, int(direct),
                                     _stream()), "fd_bn_bwd")
        if direct:
            return dx, None, None, None, None, dres, None, None, None, None, None, None
        return dx, dgamma, dbeta, None, None, dres, None, None, None, None, None, None


def batch_norm(x, gamma, beta, running_mean, running_var, residual=None, training=True,
               momentum=0.1, eps=1e-5, relu=False, stat_weight=-1.0, stats=None):
    """stats: float64 [2*C] sums already produced by conv2d(..., stats=...) for this x (training only)."""
    return BatchNormFn.apply(x, gamma, beta, running_mean, running_var, residual, bool(training),
                             float(momentum), float(eps), bool(relu), float(stat_weight),
                             stats if training else None)


# --------------------------------------------------------------------------------------------
class MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _require_cuda(x, "maxpool3x3s2")
        x = nhwc(x)
        B, C, H, W = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = empty_nhwc(B, C, Ho, Wo, x.device)
        idx = torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.uint8)
        _lib.check(_lib.load().fd_maxpool3x3s2_fwd(_p(x), _p(y), _p(idx), B, H, W, C, _stream()),
                   "fd_maxpool3x3s2_fwd")
        ctx.save_for_backward(idx)
        ctx.shape = (B, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        B, C, H, W = ctx.shape
        dy = nhwc(dy)
        dx = empty_nhwc(B, C, H, W, dy.device)
        _lib.check(_lib.load().fd_maxpool3x3s2_bwd(_p(dy), _p(idx), _p(dx), B, H, W, C, _stream()),
                   "fd_maxpool3x3s2_bwd")
        return dx


def maxpool3x3s2(x):
    return MaxPoolFn.apply(x)


# --------------------------------------------------------------------------------------------
class AssembleFn(torch.autograd.Function):
    """out = reflect_pad(cat([seg_i])) with seg_i = (a_i [+ b_i]) optionally nearest-upsampled x2."""

    @staticmethod
    def forward(ctx, pad, spec, *tensors):
        # spec: tuple of (has_b, up) per segment; tensors: a0,[b0],a1,[b1],...
        lib = _lib.load()
        segs = (_lib.Segment * len(spec))()
        keep, it, Cs = [], iter(tensors), []
        H = W = None
        for i, (has_b, up) in enumerate(spec):
            a = nhwc(next(it))
            _require_cuda(a, "assemble")
            b = nhwc(next(it)) if has_b else None
            if b is not None and b.shape != a.shape:
                raise RuntimeError("assemble: addend shape mismatch %s vs %s" % (a.shape, b.shape))
            h, w = a.shape[2] * (2 if up else 1), a.shape[3] * (2 if up else 1)
            if H is None:
                H, W = h, w
            elif (H, W) != (h, w):
                raise RuntimeError("assemble: segment %d is %dx%d, expected %dx%d" % (i, h, w, H, W))
            segs[i].a, segs[i].b = a.data_ptr(), (b.data_ptr() if b is not None else None)
            segs[i].C, segs[i].up = a.shape[1], int(up)
            Cs.append(a.shape[1])
            keep += [a, b]
        B = keep[0].shape[0]
        out = empty_nhwc(B, sum(Cs), H + 2 * pad, W + 2 * pad, keep[0].device)
        _lib.check(lib.fd_assemble_fwd(segs, len(spec), _p(out), B, H, W, pad, _stream()),
                   "fd_assemble_fwd")
        ctx.cfg = (pad, spec, Cs, B, H, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        pad, spec, Cs, B, H, W = ctx.cfg
        dout = nhwc(dout)
        n = len(spec)
        ptrs, ptrs2 = (c_void_p * n)(), (c_void_p * n)()
        Carr = (c_int * n)(*Cs)
        uarr = (c_int * n)(*[int(u) for _, u in spec])
        grads: List[Optional[torch.Tensor]] = []
        need = ctx.needs_input_grad[2:]
        k = 0
        dsegs, dsegs2 = [], []
        for i, (has_b, up) in enumerate(spec):
            wants = need[k] or (has_b and need[k + 1])
            u = 2 if up else 1
            d = empty_nhwc(B, Cs[i], H // u, W // u, dout.device) if wants else None
            # an `a + b` segment gets one gradient tensor PER operand, both written by the kernel.  The operands
            # come from trunks on different streams; a single tensor handed to both lets the autograd engine
            # accumulate the other contribution into it in place (use_count == 1 once the first consumer is done)
            # on one stream while the other stream's kernels still read it -- nothing orders the two inside a
            # captured graph (seen as an occasional wrong encoder gradient in graph replays)
            d2 = empty_nhwc(B, Cs[i], H // u, W // u, dout.device) if (has_b and need[k] and need[k + 1]) else None
            ptrs[i] = d.data_ptr() if d is not None else None
            ptrs2[i] = d2.data_ptr() if d2 is not None else None
            dsegs.append(d)
            dsegs2.append(d2)
            k += 2 if has_b else 1
        _lib.check(lib.fd_assemble_bwd2(_p(dout), ptrs, ptrs2, Carr, uarr, n, B, H, W, pad, _stream()),
                   "fd_assemble_bwd2")
        k = 0
        for i, (has_b, up) in enumerate(spec):
            grads.append(dsegs[i] if need[k] else None)
            if has_b:
                grads.append((dsegs2[i] if dsegs2[i] is not None else dsegs[i]) if need[k + 1] else None)
            k += 2 if has_b else 1
        return (None, None) + tuple(grads)


def assemble(segments: Sequence, pad: int = 1):
    """segments: sequence of (a, b_or_None, upsample_bool)."""
    spec, flat = [], []
    for a, b, up in segments:
        spec.append((b is not None, bool(up)))
        flat.append(a)
        if b is not None:
            flat.append(b)
    return AssembleFn.apply(int(pad), tuple(spec), *flat)


# --------------------------------------------------------------------------------------------
class AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        _require_cuda(a, "add")
        a, b = nhwc(a), nhwc(b)
        out = torch.empty_like(a)
        _lib.check(_lib.load().fd_add(_p(a), _p(b), _p(out), a.numel(), _stream()), "fd_add")
        return out

    @staticmethod
    def backward(ctx, g):
        # one tensor per operand: the operands' branches run on different streams (see AssembleFn.backward)
        return g, (g.clone() if all(ctx.needs_input_grad) else g)


def add(a, b):
    return AddFn.apply(a, b)


def add_relu(a, b):
    """relu(a + b), inference only (no autograd)."""
    _require_cuda(a, "add_relu")
    a, b = nhwc(a), nhwc(b)
    out = torch.empty_like(a)
    _lib.check(_lib.load().fd_add_relu(_p(a), _p(b), _p(out), a.numel(), _stream()), "fd_add_relu")
    return out


class MeanHWFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        _require_cuda(x, "mean_hw")
        x = nhwc(x)
        B, C, H, W = x.shape
        y = torch.empty((B, C), device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().fd_mean_hw_fwd(_p(x), _p(y), B, H * W, C, scale, _stream()),
                   "fd_mean_hw_fwd")
        ctx.cfg = (B, C, H, W, scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, C, H, W, scale = ctx.cfg
        dy = dy.contiguous()
        dx = empty_nhwc(B, C, H, W, dy.device)
        _lib.check(_lib.load().fd_mean_hw_bwd(_p(dy), _p(dx), B, H * W, C, scale, _stream()),
                   "fd_mean_hw_bwd")
        return dx, None


def mean_hw(x, scale=1.0):
    return MeanHWFn.apply(x, float(scale))


# --------------------------------------------------------------------------------------------
LOSS_NAMES = ["loss/0", "loss/1", "loss/2", "loss/3", "loss/si_loss0", "loss/si_loss1",
              "loss/si_loss2", "loss/si_loss3", "loss"]


class PhotoLossFn(torch.autograd.Function):
    """Fused generate_images_pred + compute_losses (trainer.py:425-596).

    Differentiable inputs: the four disparities and the two poses.  Output: losses[9] in the
    order of LOSS_NAMES; only losses[8] ("loss") carries gradient (that is the only entry the
    reference back-propagates, trainer.py:244-245)."""

    @staticmethod
    def forward(ctx, d0, d1, d2, d3, T0, T1, static: Dict, opts: Dict, outs: Optional[Dict]):
        lib = _lib.load()
        _require_cuda(d0, "photoloss")
        disps = [d.contiguous() for d in (d0, d1, d2, d3)]
        T0, T1 = T0.contiguous(), T1.contiguous()
        B, _, H, W = disps[0].shape
        dev = d0.device
        desc = _lib.PhotolossDesc()
        desc.B, desc.H, desc.W = B, H, W
        keep = []
        for i, f in enumerate((0, -1, 1)):
            for s in range(4):
                t = static["color"].get((f, s))
                if t is not None:
                    t = t.contiguous()
                    keep.append(t)
                    desc.color[i][s] = t.data_ptr()
        for s in range(4):
            if disps[s].shape != (B, 1, H >> s, W >> s):
                raise RuntimeError("photoloss: disp %d has shape %s" % (s, tuple(disps[s].shape)))
            desc.disp[s] = disps[s].data_ptr()
            n = static["noise"][s].contiguous()
            keep.append(n)
            desc.noise[s] = n.data_ptr()
        K, invK, beam = static["K"].contiguous(), static["inv_K"].contiguous(), static["beam"].contiguous()
        keep += [K, invK, beam]
        desc.K, desc.inv_K, desc.beam = K.data_ptr(), invK.data_ptr(), beam.data_ptr()
        desc.T[0], desc.T[1] = T0.data_ptr(), T1.data_ptr()
        desc.min_depth, desc.max_depth = opts.get("min_depth", 0.1), opts.get("max_depth", 100.0)
        desc.smoothness = opts.get("smoothness", 1e-3)
        desc.si_thresh, desc.si_var = opts.get("si_thresh", 2.0), opts.get("si_var", 0.3)
        # si-loss term: trainer.py:577-589 defaults; "use_si" False switches it off, "si_scales" is a bit
        # mask of scales, the other keys select the refiner's GDC-clone variant (refiner.py:678-688)
        desc.si_scales = int(opts.get("si_scales", 0xF)) if opts.get("use_si", True) else 0
        desc.si_pred_mul, desc.si_tgt_mul = opts.get("si_pred_mul", 26.0), opts.get("si_tgt_mul", 100.0)
        desc.si_lo, desc.si_weight = opts.get("si_lo", 1.0), opts.get("si_weight", 0.1)
        sel = torch.empty((4, B, H, W), device=dev, dtype=torch.uint8)
        desc.sel = sel.data_ptr()
        if outs is not None:
            for s in range(4):
                outs[("depth", 0, s)] = torch.empty((B, 1, H, W), device=dev)
                desc.out_depth[s] = outs[("depth", 0, s)].data_ptr()
                outs["to_optimise/%d" % s] = torch.empty((B, H, W), device=dev)
                desc.out_to_optimise[s] = outs["to_optimise/%d" % s].data_ptr()
                for j, f in enumerate((-1, 1)):
                    outs[("color", f, s)] = torch.empty((B, 3, H, W), device=dev)
                    desc.out_color[s][j] = outs[("color", f, s)].data_ptr()
            outs["sel"] = sel
        ws = torch.empty(lib.fd_photoloss_workspace_bytes(B, H, W) // 4, device=dev, dtype=torch.float32)
        losses = torch.empty(9, device=dev, dtype=torch.float32)
        with _timed("photoloss_fwd", 213.25 * B * H * W):
            _lib.check(lib.fd_photoloss_fwd(ctypes.byref(desc), _p(losses), _p(ws), _stream()),
                       "fd_photoloss_fwd")
        # the forward-only outputs must not be rewritten by the backward's descriptor
        for s in range(4):
            desc.out_depth[s] = None
            desc.out_to_optimise[s] = None
            desc.out_color[s][0] = None
            desc.out_color[s][1] = None
        ctx.desc, ctx.keep, ctx.ws, ctx.sel = desc, keep + disps + [T0, T1], ws, sel
        ctx.shape = (B, H, W)
        return losses

    @staticmethod
    def backward(ctx, glosses):
        lib = _lib.load()
        B, H, W = ctx.shape
        dev = glosses.device
        g = glosses[8:9].contiguous()
        gd = [torch.empty((B, 1, H >> s, W >> s), device=dev, dtype=torch.float32) for s in range(4)]
        gT0 = torch.empty((B, 4, 4), device=dev, dtype=torch.float32)
        gT1 = torch.empty((B, 4, 4), device=dev, dtype=torch.float32)
        arr = (c_void_p * 4)(*[t.data_ptr() for t in gd])
        with _timed("photoloss_bwd", 218.56 * B * H * W):
            _lib.check(lib.fd_photoloss_bwd(ctypes.byref(ctx.desc), _p(g), ctypes.byref(arr), _p(gT0),
                                            _p(gT1), _p(ctx.ws), _stream()), "fd_photoloss_bwd")
        return gd[0], gd[1], gd[2], gd[3], gT0, gT1, None, None, None


def photoloss(disps: Sequence[torch.Tensor], T_m1: torch.Tensor, T_p1: torch.Tensor, static: Dict,
              opts: Optional[Dict] = None, outs: Optional[Dict] = None) -> torch.Tensor:
    return PhotoLossFn.apply(disps[0], disps[1], disps[2], disps[3], T_m1, T_p1, static, opts or {},
                             outs)


# --------------------------------------------------------------------------------------------
# Unfused drop-in operators (csrc/geometry.cu): the layers.* modules and the layers.F proxy call these, so
# that an UNPATCHED reference driver still runs its loss chain on this library's kernels.
def _nchw(t):
    return t.contiguous()


class UpsampleBilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H, W):
        _require_cuda(x, "upsample_bilinear")
        x = _nchw(x)
        B, C, h, w = x.shape
        y = torch.empty((B, C, H, W), device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().fd_upsample_bilinear_fwd(_p(x), _p(y), B * C, h, w, H, W, _stream()),
                   "fd_upsample_bilinear_fwd")
        ctx.shape = (B, C, h, w, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, C, h, w, H, W = ctx.shape
        dy = _nchw(dy)
        dx = torch.empty((B, C, h, w), device=dy.device, dtype=torch.float32)
        _lib.check(_lib.load().fd_upsample_bilinear_bwd(_p(dy), _p(dx), B * C, h, w, H, W, _stream()),
                   "fd_upsample_bilinear_bwd")
        return dx, None, None


def upsample_bilinear(x, H, W):
    """F.interpolate(x, [H, W], mode="bilinear", align_corners=False)."""
    return UpsampleBilinearFn.apply(x, int(H), int(W))


class BackprojectFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, inv_K, H, W):
        _require_cuda(depth, "backproject")
        depth, inv_K = _nchw(depth), _nchw(inv_K)
        B = depth.shape[0]
        if depth.numel() != B * H * W or inv_K.shape != (B, 4, 4):
            # the reference fails with a broadcasting RuntimeError when the batch differs from the ctor's
            raise RuntimeError("BackprojectDepth: depth %s / inv_K %s do not match batch %d x %d x %d"
                               % (tuple(depth.shape), tuple(inv_K.shape), B, H, W))
        cam = torch.empty((B, 4, H * W), device=depth.device, dtype=torch.float32)
        _lib.check(_lib.load().fd_backproject_fwd(_p(depth), _p(inv_K), _p(cam), B, H, W, _stream()),
                   "fd_backproject_fwd")
        ctx.save_for_backward(inv_K)
        ctx.shape = (tuple(depth.shape), B, H, W)
        return cam

    @staticmethod
    def backward(ctx, dcam):
        (inv_K,) = ctx.saved_tensors
        shape, B, H, W = ctx.shape
        dcam = _nchw(dcam)
        dd = torch.empty(shape, device=dcam.device, dtype=torch.float32)
        _lib.check(_lib.load().fd_backproject_bwd(_p(dcam), _p(inv_K), _p(dd), B, H, W, _stream()),
                   "fd_backproject_bwd")
        return dd, None, None, None


def backproject(depth, inv_K, H, W):
    return BackprojectFn.apply(depth, inv_K, int(H), int(W))


class Project3DFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, K, T, H, W, eps):
        _require_cuda(points, "project3d")
        points, K, T = _nchw(points), _nchw(K), _nchw(T)
        B = points.shape[0]
        if points.shape != (B, 4, H * W):
            raise RuntimeError("Project3D: points %s do not match %d x 4 x %d" % (tuple(points.shape), B, H * W))
        grid = torch.empty((B, H, W, 2), device=points.device, dtype=torch.float32)
        _lib.check(_lib.load().fd_project3d_fwd(_p(points), _p(K), _p(T), _p(grid), B, H, W, eps, _stream()),
                   "fd_project3d_fwd")
        ctx.save_for_backward(points, K, T)
        ctx.cfg = (B, H, W, eps)
        return grid

    @staticmethod
    def backward(ctx, dgrid):
        points, K, T = ctx.saved_tensors
        B, H, W, eps = ctx.cfg
        dgrid = _nchw(dgrid)
        dp = torch.empty_like(points) if ctx.needs_input_grad[0] else None
        dT = torch.empty_like(T) if ctx.needs_input_grad[2] else None
        ws = torch.empty(B * 12, device=points.device, dtype=torch.float32)
        _lib.check(_lib.load().fd_project3d_bwd(_p(points), _p(K), _p(T), _p(dgrid), _p(dp), _p(dT), B, H, W,
                                                eps, _p(ws), _stream()), "fd_project3d_bwd")
        return dp, None, dT, None, None, None


def project3d(points, K, T, H, W, eps=1e-7):
    if K.requires_grad:
        raise NotImplementedError("project3d: no gradient wrt the intrinsics K")
    return Project3DFn.apply(points, K, T, int(H), int(W), float(eps))


class GridSampleBorderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, grid):
        _require_cuda(img, "grid_sample")
        img, grid = _nchw(img), _nchw(grid)
        B, C, H, W = img.shape
        if grid.dim() != 4 or grid.shape[0] != B or grid.shape[3] != 2:
            raise RuntimeError("grid_sample: grid %s does not match input %s" % (tuple(grid.shape), tuple(img.shape)))
        Ho, Wo = grid.shape[1], grid.shape[2]
        out = torch.empty((B, C, Ho, Wo), device=img.device, dtype=torch.float32)
        _lib.check(_lib.load().fd_grid_sample_border_fwd(_p(img), _p(grid), _p(out), B, C, H, W, Ho, Wo, _stream()),
                   "fd_grid_sample_border_fwd")
        ctx.save_for_backward(img, grid)
        return out

    @staticmethod
    def backward(ctx, dout):
        img, grid = ctx.saved_tensors
        B, C, H, W = img.shape
        Ho, Wo = grid.shape[1], grid.shape[2]
        dout = _nchw(dout)
        dimg = torch.empty_like(img) if ctx.needs_input_grad[0] else None
        dgrid = torch.empty_like(grid) if ctx.needs_input_grad[1] else None
        _lib.check(_lib.load().fd_grid_sample_border_bwd(_p(img), _p(grid), _p(dout), _p(dgrid), _p(dimg), B, C,
                                                         H, W, Ho, Wo, _stream()), "fd_grid_sample_border_bwd")
        return dimg, dgrid


def grid_sample_border(img, grid):
    """F.grid_sample(img, grid, mode="bilinear", padding_mode="border", align_corners=False)."""
    return GridSampleBorderFn.apply(img, grid)


class SSIMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        _require_cuda(x, "ssim")
        x, y = _nchw(x), _nchw(y)
        if x.shape != y.shape or x.dim() != 4:
            raise RuntimeError("ssim: shapes %s / %s" % (tuple(x.shape), tuple(y.shape)))
        B, C, H, W = x.shape
        out = torch.empty_like(x)
        _lib.check(_lib.load().fd_ssim_fwd(_p(x), _p(y), _p(out), B * C, H, W, _stream()), "fd_ssim_fwd")
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y = ctx.saved_tensors
        B, C, H, W = x.shape
        dout = _nchw(dout)
        lib = _lib.load()
        ws = torch.empty(3 * x.numel(), device=x.device, dtype=torch.float32)
        dx = dy = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _lib.check(lib.fd_ssim_bwd(_p(x), _p(y), _p(dout), _p(dx), B * C, H, W, _p(ws), _stream()), "fd_ssim_bwd")
        if ctx.needs_input_grad[1]:
            dy = torch.empty_like(y)
            _lib.check(lib.fd_ssim_bwd(_p(y), _p(x), _p(dout), _p(dy), B * C, H, W, _p(ws), _stream()), "fd_ssim_bwd")
        return dx, dy


def ssim(x, y):
    return SSIMFn.apply(x, y)


class PoseMatrixFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, axisangle, translation, invert):
        _require_cuda(axisangle, "pose_matrix")
        aa = axisangle.contiguous().view(-1, 3)
        tr = translation.contiguous().view(-1, 3)
        B = aa.shape[0]
        if tr.shape[0] != B:
            raise RuntimeError("transformation_from_parameters: %s vs %s" % (tuple(axisangle.shape),
                                                                             tuple(translation.shape)))
        M = torch.empty((B, 4, 4), device=aa.device, dtype=torch.float32)
        _lib.check(_lib.load().fd_pose_matrix_fwd(_p(aa), _p(tr), int(invert), _p(M), B, _stream()),
                   "fd_pose_matrix_fwd")
        ctx.save_for_backward(aa, tr)
        ctx.cfg = (int(invert), axisangle.shape, translation.shape)
        return M

    @staticmethod
    def backward(ctx, dM):
        aa, tr = ctx.saved_tensors
        invert, sa, st = ctx.cfg
        dM = dM.contiguous()
        daa, dtr = torch.empty_like(aa), torch.empty_like(tr)
        _lib.check(_lib.load().fd_pose_matrix_bwd(_p(aa), _p(tr), invert, _p(dM), _p(daa), _p(dtr), aa.shape[0],
                                                  _stream()), "fd_pose_matrix_bwd")
        return daa.view(sa), dtr.view(st), None


def pose_matrix(axisangle, translation, invert=False):
    """layers.transformation_from_parameters: [B,1,3] x 2 -> [B,4,4], one kernel each way."""
    return PoseMatrixFn.apply(axisangle, translation, bool(invert))


def cat_xy(depth, inv_K, H, W):
    """layers.Cat_xy.forward.  No gradient (the refiner calls it under no_grad, refiner.py:306-346)."""
    _require_cuda(depth, "cat_xy")
    if depth.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("Cat_xy: no gradient wrt depth (the reference uses it under no_grad)")
    depth, inv_K = _nchw(depth), _nchw(inv_K)
    B = depth.shape[0]
    if depth.numel() != B * H * W or inv_K.shape != (B, 4, 4):
        raise RuntimeError("Cat_xy: depth %s / inv_K %s do not match batch %d x %d x %d"
                           % (tuple(depth.shape), tuple(inv_K.shape), B, H, W))
    out = torch.empty((B, 3, H, W), device=depth.device, dtype=torch.float32)
    _lib.check(_lib.load().fd_cat_xy(_p(depth), _p(inv_K), _p(out), B, H, W, _stream()), "fd_cat_xy")
    return out


# --------------------------------------------------------------------------------------------
def refine_pack(disp0, beam, two_cha, inv_Ks, crop=(78, 190, 23, 617), min_depth=0.1, max_depth=100.0):
    """The stage-2 pseudo-3D maps of refiner.py:316-346 (fd_refine_pack).  Returns ([B,6,h,w] x 4 stored
    channels-last, ratios[4]).  No gradient: the reference builds these under no_grad."""
    _require_cuda(disp0, "refine_pack")
    lib = _lib.load()
    with torch.no_grad():
        disp0, beam, two_cha = _nchw(disp0.detach()), _nchw(beam), _nchw(two_cha)
        B, _, H, W = disp0.shape
        if beam.numel() != B * H * W or two_cha.shape != (B, 2, H, W):
            raise RuntimeError("refine_pack: beam %s / two_cha %s do not match disp %s"
                               % (tuple(beam.shape), tuple(two_cha.shape), tuple(disp0.shape)))
        iks = [_nchw(k) for k in inv_Ks]
        outs = [empty_nhwc(B, 6, H >> s, W >> s, disp0.device) for s in range(4)]
        ratios = torch.empty(4, device=disp0.device, dtype=torch.float32)
        ws = torch.empty(lib.fd_refine_pack_workspace_bytes(B, H, W) // 4, device=disp0.device, dtype=torch.float32)
        ik_arr = (c_void_p * 4)(*[t.data_ptr() for t in iks])
        out_arr = (c_void_p * 4)(*[t.data_ptr() for t in outs])
        _lib.check(lib.fd_refine_pack(_p(disp0), _p(beam), _p(two_cha), ctypes.byref(ik_arr), B, H, W,
                                      int(crop[0]), int(crop[1]), int(crop[2]), int(crop[3]), min_depth, max_depth,
                                      ctypes.byref(out_arr), _p(ratios), _p(ws), _stream()), "fd_refine_pack")
    return outs, ratios


DEPTH_METRIC_NAMES = ("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3")


def depth_errors(gt, pred, window=None, mask_lo=0.0, mask_hi=float("inf"), pred_is_disp=False,
                 pre_clamp=(-float("inf"), float("inf")), median_scaling=True, numpy_median=False,
                 clamp=(1e-3, 80.0)):
    """The 7 depth metrics (+ count, ratio) of layers.compute_depth_errors behind the reference's masking /
    median scaling / clamping (fd_depth_errors): a [9] device tensor.  gt and pred have the same size."""
    _require_cuda(gt, "depth_errors")
    lib = _lib.load()
    gt, pred = _nchw(gt.detach()), _nchw(pred.detach())
    if gt.numel() != pred.numel():
        raise RuntimeError("depth_errors: gt %s and prediction %s differ in size" % (tuple(gt.shape), tuple(pred.shape)))
    H, W = gt.shape[-2:]
    B = gt.numel() // (H * W)
    y0, y1, x0, x1 = window if window is not None else (0, H, 0, W)
    out = torch.empty(9, device=gt.device, dtype=torch.float32)
    ws = torch.empty(lib.fd_depth_errors_workspace_bytes(B, H, W) // 4, device=gt.device, dtype=torch.float32)
    _lib.check(lib.fd_depth_errors(_p(gt), _p(pred), B, H, W, int(y0), int(y1), int(x0), int(x1), float(mask_lo),
                                   float(mask_hi), int(pred_is_disp), float(pre_clamp[0]), float(pre_clamp[1]),
                                   int(median_scaling), int(numpy_median), float(clamp[0]), float(clamp[1]),
                                   _p(out), _p(ws), _stream()), "fd_depth_errors")
    return out


def masked_median(x, mask_src, window=None, scale=1.0):
    """torch.median((x * scale)[mask_src > 0 inside window]) -- lower median, 0-dim tensor (refiner.py:332)."""
    _require_cuda(x, "masked_median")
    lib = _lib.load()
    x, mask_src = _nchw(x.detach()), _nchw(mask_src)
    if x.shape != mask_src.shape:
        raise RuntimeError("masked_median: shapes %s / %s" % (tuple(x.shape), tuple(mask_src.shape)))
    H, W = x.shape[-2:]
    B = x.numel() // (H * W)
    y0, y1, x0, x1 = window if window is not None else (0, H, 0, W)
    out = torch.empty(1, device=x.device, dtype=torch.float32)
    ws = torch.empty(lib.fd_masked_median_workspace_bytes(B, H, W) // 4, device=x.device, dtype=torch.float32)
    _lib.check(lib.fd_masked_median(_p(x), _p(mask_src), B, H, W, int(y0), int(y1), int(x0), int(x1), float(scale),
                                    _p(out), _p(ws), _stream()), "fd_masked_median")
    return out[0]


# --------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, state, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """In-place Adam on flat fp32 buffers; `state` is a 4-word int32 device tensor (step counter, ...,
    learning rate as float bits in word 3 -- used when lr < 0)."""
    _lib.check(_lib.load().fd_adam_step(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps,
                                        _p(state), grad_scale, _stream()), "fd_adam_step")
