"""Functional host-side wrappers: torch tensors in, C-ABI calls on the current CUDA stream.

PyTorch is plumbing here (device memory, streams); every byte of arithmetic happens in
libsmaat_b200.so.  Tensors must be fp32 CUDA tensors, NCHW, dense in (C?, H, W) -- a
batch stride larger than C*H*W is allowed where the C ABI takes a ``bstride``.
"""
from __future__ import annotations

import os

import torch

from . import _lib

PW_MODES = {"fp32": 0, "tf32": 1, "tf32x3": 2}
_pw_mode = os.environ.get("SMAAT_PW_MODE", "tf32x3")
assert _pw_mode in PW_MODES, f"SMAAT_PW_MODE must be one of {list(PW_MODES)}"


def set_pointwise_mode(mode: str) -> None:
    """'tf32x3' (default; tcgen05 3xTF32 split, fp32-grade), 'tf32' (tcgen05 single pass; what
    cuDNN's allow_tf32=True default gives the reference on a GPU), 'fp32' (CUDA-core exact)."""
    global _pw_mode
    if mode not in PW_MODES:
        raise ValueError(f"pointwise mode must be one of {list(PW_MODES)}")
    _pw_mode = mode


def get_pointwise_mode() -> str:
    return _pw_mode


# ---- generation counter of "weights written behind torch's back" (see modules._versions) ----
_weights_gen = 0


def weights_generation() -> int:
    return _weights_gen


def bump_weights_generation() -> None:
    """Call after parameters / buffers were (or may have been) written through raw pointers or a CUDA-graph replay:
    invalidates every cache derived from them (folded BatchNorm affine, tf32 hi/lo weight splits)."""
    global _weights_gen
    _weights_gen += 1


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ---- optional per-launch timing (bench.py roofline leg): CUDA events on the launching stream ----
_prof = None


class profile:
    """Context manager: records (kernel name, algorithmic bytes, flops, start/end CUDA events) for every
    C-ABI launch made inside it.  ``summary()`` synchronises and aggregates per kernel name."""

    def __enter__(self):
        global _prof
        self.records = []
        _prof = self
        return self

    def __exit__(self, *exc):
        global _prof
        _prof = None
        return False

    def summary(self, by_shape=False):
        torch.cuda.synchronize()
        agg = {}
        for name, nbytes, flops, e0, e1 in self.records:
            key = name if by_shape else name.split("[")[0]
            a = agg.setdefault(key, {"launches": 0, "ms": 0.0, "bytes": 0, "flops": 0})
            a["launches"] += 1
            a["ms"] += e0.elapsed_time(e1)
            a["bytes"] += nbytes
            a["flops"] += flops
        return agg


def _call(name, nbytes, flops, fn, *args):
    """Invoke one C-ABI entry point on the current stream (optionally bracketed by timing events)."""
    if _prof is None:
        _lib.check(fn(*args), name.split("[")[0])
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(fn(*args), name)
    e1.record()
    _prof.records.append((name, int(nbytes), int(flops), e0, e1))


def _ptr(t):
    return None if t is None else t.data_ptr()


def _req(t, name, ndim=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError(f"smaat_unet_b200: {name} must be a float32 CUDA tensor "
                           f"(got {type(t).__name__} {getattr(t, 'dtype', None)} {getattr(t, 'device', None)}); "
                           "there is no CPU fallback")
    if ndim is not None and t.dim() != ndim:
        raise RuntimeError(f"smaat_unet_b200: {name} must be {ndim}-D, got shape {tuple(t.shape)}")
    return t


def _dense(t, name):
    _req(t, name)
    return t if t.is_contiguous() else t.contiguous()


def _nchw_bstride(t, name):
    """Accept NCHW tensors that are dense in (C,H,W); return (tensor, batch stride in elements)."""
    _req(t, name, 4)
    B, Cc, H, W = t.shape
    st = t.stride()
    if (st[3] == 1 or W == 1) and (st[2] == W or H == 1) and (st[1] == H * W or Cc == 1) and (B == 1 or st[0] >= Cc * H * W):
        return t, (st[0] if B > 1 else Cc * H * W)
    t = t.contiguous()
    return t, Cc * H * W


# ---------------------------------------------------------------------------------------------
def dw3x3(x, weight, bias, k, x1=None, in_scale=None, in_shift=None, loader=0):
    """Depthwise 3x3/pad 1 over the virtual concat [x, x1] (layers.py:38-44; parts_ds.py:85)."""
    x, bs0 = _nchw_bstride(x, "x")
    B, C0, H, W = x.shape
    C1, bs1 = 0, 0
    if x1 is not None:
        x1, bs1 = _nchw_bstride(x1, "x1")
        assert x1.shape[0] == B and x1.shape[2:] == x.shape[2:], "concat inputs must agree in B, H, W"
        C1 = x1.shape[1]
    w = _dense(weight, "depthwise.weight")
    assert w.numel() == k * (C0 + C1) * 9, f"depthwise weight {tuple(w.shape)} does not match k*(C0+C1)={k * (C0 + C1)}"
    y = torch.empty((B, k * (C0 + C1), H, W), device=x.device, dtype=torch.float32)
    _call(f"smaat_dw3x3_fwd[C{C0 + C1}_S{H}]", 4 * B * H * W * (C0 + C1) * (1 + k), 18 * B * H * W * k * (C0 + C1), _lib.load().smaat_dw3x3_fwd, _ptr(x), C0, bs0, _ptr(x1), C1, bs1, _ptr(w), _ptr(bias), _ptr(in_scale), _ptr(in_shift),
                                   _ptr(y), B, H, W, k, loader, _stream())
    return y


def tc_eligible(x, w2d) -> bool:
    K, P = x.shape[1], x.shape[2] * x.shape[3]
    return bool(_lib.load().smaat_pw1x1_tc_eligible(_ptr(x), _ptr(w2d), K, w2d.shape[0], P))


def split_tf32(w):
    w = _dense(w, "w")
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    _call("smaat_split_tf32", 12 * w.numel(), 0, _lib.load().smaat_split_tf32, _ptr(w), _ptr(hi), _ptr(lo), w.numel(), _stream())
    return hi, lo


def pw1x1(x, weight, scale, shift, relu, mode=None, w_split=None, stats=None, out=None):
    """Pointwise 1x1 + per-channel affine (+ReLU) (layers.py:45,49 + parts_ds.py:25-26).

    weight: (Cout, K[,1,1]).  mode None = module-level default.  w_split = cached (hi, lo) for
    'tf32x3'.  Falls to the exact CUDA-core kernel for shapes the tcgen05 path does not take.
    """
    x = _dense(x, "x")
    B, K, H, W = x.shape
    w2d = _dense(weight, "pointwise.weight").view(weight.shape[0], -1)
    Cout = w2d.shape[0]
    assert w2d.shape[1] == K, f"pointwise weight {tuple(weight.shape)} does not match K={K}"
    P = H * W
    if out is None:
        out = torch.empty((B, Cout, H, W), device=x.device, dtype=torch.float32)
        ybs = Cout * P
    else:
        out, ybs = _nchw_bstride(out, "out")
    mode = mode or _pw_mode
    m = PW_MODES[mode]
    wlo = None
    if m != 0 and not tc_eligible(x, w2d):
        m = 0
    if m == 2:
        hi, wlo = w_split if w_split is not None else split_tf32(w2d)
        w2d = hi
    _call(f"smaat_pw1x1_fwd[K{K}_N{Cout}_P{P}]", 4 * B * P * (K + Cout) + 4 * K * Cout, 2 * B * P * K * Cout, _lib.load().smaat_pw1x1_fwd, _ptr(x), _ptr(w2d), _ptr(wlo), _ptr(scale), _ptr(shift), _ptr(out), ybs, _ptr(stats),
                                           B, K, Cout, P, int(bool(relu)), m, _stream())
    return out


_fuse_ds = os.environ.get("SMAAT_FUSE_DS", "1") != "0"


def set_fused_dsconv(enabled: bool) -> None:
    """Enable/disable the fused depthwise->pointwise kernel (default on; off = dw3x3 + pw1x1 kernels)."""
    global _fuse_ds
    _fuse_ds = bool(enabled)


def set_dsconv_impl(impl: str) -> None:
    """Which fused DS-conv kernel runs: 'auto' (TMEM-operand kernel where it applies, else the shared-memory-operand one),
    'smem' (round-1 kernel only) or 'tmem' (TMEM-operand kernel only; other shapes fall back to dw3x3 + pw1x1)."""
    _lib.check(_lib.load().smaat_set_dsconv_impl({"auto": 0, "smem": 1, "tmem": 2}[impl]), "smaat_set_dsconv_impl")


def dsconv_takes(x, x1, pw_weight, k, mode=None, stats=False) -> bool:
    """True when ``dsconv`` would run its fused kernel on these inputs (smaat_dsconv_eligible + the arithmetic mode)."""
    mode = mode or _pw_mode
    if not _fuse_ds or PW_MODES[mode] == 0:
        return False
    x, bs0 = _nchw_bstride(x, "x")
    C1, bs1 = 0, 0
    if x1 is not None:
        x1, bs1 = _nchw_bstride(x1, "x1")
        C1 = x1.shape[1]
    w2d = _dense(pw_weight, "pointwise.weight").view(pw_weight.shape[0], -1)
    return bool(_lib.load().smaat_dsconv_eligible2(_ptr(x), x.shape[1], bs0, _ptr(x1), C1, bs1, _ptr(w2d), x.shape[2], x.shape[3], k, w2d.shape[0], int(bool(stats))))


def dsconv(x, dw_weight, dw_bias, k, pw_weight, scale, shift, relu, x1=None, mode=None, w_split=None, stats=None, outconv=None):
    """Fused DepthwiseSeparableConv (layers.py:47-50) + affine (+ReLU); returns None when the fused kernel
    does not take this shape/mode (caller then runs dw3x3 + pw1x1).  ``outconv=(weight (1, Cout[,1,1]), bias or None)``
    appends the 1-class OutConv in the epilogue and returns the (B, 1, H, W) logits instead of the activation."""
    mode = mode or _pw_mode
    if not _fuse_ds or PW_MODES[mode] == 0:
        return None
    x, bs0 = _nchw_bstride(x, "x")
    B, C0, H, W = x.shape
    C1, bs1 = 0, 0
    if x1 is not None:
        x1, bs1 = _nchw_bstride(x1, "x1")
        C1 = x1.shape[1]
    w2d = _dense(pw_weight, "pointwise.weight").view(pw_weight.shape[0], -1)
    Cout, K = w2d.shape
    assert K == k * (C0 + C1), f"pointwise weight {tuple(pw_weight.shape)} does not match k*Cin={k * (C0 + C1)}"
    lib = _lib.load()
    if not lib.smaat_dsconv_eligible2(_ptr(x), C0, bs0, _ptr(x1), C1, bs1, _ptr(w2d), H, W, k, Cout, int(stats is not None)):
        return None
    wlo = None
    if PW_MODES[mode] == 2:
        w2d, wlo = w_split if w_split is not None else split_tf32(w2d)
    dw_w = _dense(dw_weight, "depthwise.weight")
    Cin = C0 + C1
    if outconv is not None:
        ow, ob = outconv
        assert ow.numel() == Cout and stats is None, "fused OutConv: one class over the block's Cout channels, no batch statistics"
        logits = torch.empty((B, 1, H, W), device=x.device, dtype=torch.float32)
        _call(f"smaat_dsconv_outconv_fwd[C{Cin}_N{Cout}_S{H}]", 4 * B * H * W * (Cin + 1) + 4 * K * Cout, 2 * B * H * W * (K * (Cout + 9) + Cout),
              lib.smaat_dsconv_outconv_fwd, _ptr(x), C0, bs0, _ptr(x1), C1, bs1, _ptr(dw_w), _ptr(dw_bias), _ptr(w2d), _ptr(wlo), _ptr(scale),
              _ptr(shift), _ptr(_dense(ow, "outconv.weight")), _ptr(ob), _ptr(logits), B, H, W, k, Cout, int(bool(relu)), PW_MODES[mode], _stream())
        return logits
    y = torch.empty((B, Cout, H, W), device=x.device, dtype=torch.float32)
    _call(f"smaat_dsconv_fwd[C{Cin}_N{Cout}_S{H}]", 4 * B * H * W * (Cin + Cout) + 4 * K * Cout, 2 * B * H * W * K * (Cout + 9),
          lib.smaat_dsconv_fwd, _ptr(x), C0, bs0, _ptr(x1), C1, bs1, _ptr(dw_w), _ptr(dw_bias), _ptr(w2d), _ptr(wlo), _ptr(scale),
          _ptr(shift), _ptr(y), Cout * H * W, _ptr(stats), B, H, W, k, Cout, int(bool(relu)), PW_MODES[mode], _stream())
    return y


def bn_fold(gamma, beta, running_mean, running_var, conv_bias, eps):
    """Eval BatchNorm2d -> (scale, shift) for the pw epilogue (parts_ds.py:25,34)."""
    Cn = gamma.numel()
    scale = torch.empty(Cn, device=gamma.device, dtype=torch.float32)
    shift = torch.empty_like(scale)
    _call("smaat_bn_fold", 28 * Cn, 0, _lib.load().smaat_bn_fold, _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), _ptr(conv_bias),
                                         float(eps), _ptr(scale), _ptr(shift), Cn, _stream())
    return scale, shift


def new_stats(C, device):
    """Zeroed fp64 accumulators [sum(C) | sum of squares(C)] for the train-mode BN epilogues."""
    return torch.zeros(2 * C, device=device, dtype=torch.float64)


def channel_stats(x, stats=None):
    x = _dense(x, "x")
    B, Cc, H, W = x.shape
    if stats is None:
        stats = new_stats(Cc, x.device)
    _call("smaat_channel_stats", 4 * B * Cc * H * W, 0, _lib.load().smaat_channel_stats, _ptr(x), _ptr(stats), B, Cc, H * W, _stream())
    return stats


def bn_finalize(stats, count, bn, save=False):
    """Batch statistics -> (scale, shift) [, mean, invstd]; updates bn.running_mean/var in place (torch semantics)."""
    Cn = bn.num_features
    dev = stats.device
    scale = torch.empty(Cn, device=dev, dtype=torch.float32)
    shift = torch.empty_like(scale)
    mean = torch.empty_like(scale) if save else None
    invstd = torch.empty_like(scale) if save else None
    track = bn.track_running_stats and bn.running_mean is not None
    if track and bn.momentum is None:
        raise NotImplementedError("BatchNorm2d(momentum=None) (cumulative average) is not supported")
    if track:
        bump_weights_generation()      # running statistics are written by raw pointer: no _version bump
    _call("smaat_bn_finalize", 40 * Cn, 0, _lib.load().smaat_bn_finalize, _ptr(stats), float(count),
          _ptr(bn.weight.detach() if bn.weight is not None else None), _ptr(bn.bias.detach() if bn.bias is not None else None),
          float(bn.eps), float(bn.momentum if bn.momentum is not None else 0.1),
          _ptr(bn.running_mean if track else None), _ptr(bn.running_var if track else None),
          _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd),
          _ptr(bn.num_batches_tracked if (track and bn.num_batches_tracked is not None) else None), Cn, _stream())
    return (scale, shift, mean, invstd) if save else (scale, shift)


def affine_act(x, scale, shift, act):
    """y = act(scale[c]*x + shift[c]); act in {'none','relu','sigmoid'}."""
    x = _dense(x, "x")
    B, Cc, H, W = x.shape
    y = torch.empty_like(x)
    code = {"none": 0, "relu": 1, "sigmoid": 2}[act]
    _call("smaat_affine_act_fwd", 8 * B * Cc * H * W, 0, _lib.load().smaat_affine_act_fwd, _ptr(x), _ptr(scale), _ptr(shift), _ptr(y),
          B, Cc, H * W, code, _stream())
    return y


def maxpool2(x):
    """nn.MaxPool2d(2) (parts_ds.py:48)."""
    x = _dense(x, "x")
    B, Cc, H, W = x.shape
    y = torch.empty((B, Cc, H // 2, W // 2), device=x.device, dtype=torch.float32)
    _call("smaat_maxpool2_fwd", 4 * B * Cc * (H * W + (H // 2) * (W // 2)), 0, _lib.load().smaat_maxpool2_fwd, _ptr(x), _ptr(y), B * Cc, H, W, _stream())
    return y


def upsample2x_pad(x, Ho, Wo):
    """nn.Upsample(x2, bilinear, align_corners=True) + F.pad to (Ho, Wo) (parts_ds.py:64,78-81)."""
    x = _dense(x, "x")
    B, Cc, H, W = x.shape
    y = torch.empty((B, Cc, Ho, Wo), device=x.device, dtype=torch.float32)
    _call(f"smaat_upsample2x_pad_fwd[C{Cc}_S{H}]", 4 * B * Cc * (H * W + Ho * Wo), 0, _lib.load().smaat_upsample2x_pad_fwd, _ptr(x), _ptr(y), Cc * Ho * Wo, B, Cc, H, W, Ho, Wo, _stream())
    return y


def convt2x2_pack_weight(w):
    """ConvTranspose2d(k=2, s=2) weight (Cin, Cout, 2, 2) -> the pointwise GEMM's (4 Cout, Cin) matrix (csrc/convt.cu)."""
    w = _dense(w, "up.weight")
    Cin, Cout = w.shape[0], w.shape[1]
    wp = torch.empty((4 * Cout, Cin), device=w.device, dtype=torch.float32)
    _call("smaat_convt2x2_pack_weight", 8 * w.numel(), 0, _lib.load().smaat_convt2x2_pack_weight, _ptr(w), _ptr(wp), Cin, Cout, _stream())
    return wp


def pixel_shuffle2_pad(t, bias, Cout, Ho, Wo):
    """(B, 4 Cout, H, W) packed taps -> (B, Cout, Ho, Wo): 2x2 pixel shuffle + bias + F.pad frame (parts_ds.py:76-81)."""
    t = _dense(t, "t")
    B, C4, H, W = t.shape
    assert C4 == 4 * Cout
    y = torch.empty((B, Cout, Ho, Wo), device=t.device, dtype=torch.float32)
    _call("smaat_pixel_shuffle2_pad_fwd", 4 * B * Cout * (4 * H * W + Ho * Wo), 0, _lib.load().smaat_pixel_shuffle2_pad_fwd, _ptr(t), _ptr(bias), _ptr(y),
          Cout * Ho * Wo, B, Cout, H, W, Ho, Wo, _stream())
    return y


def cbam_pool(x):
    x = _dense(x, "x")
    B, Cc, H, W = x.shape
    avg = torch.empty((B, Cc), device=x.device, dtype=torch.float32)
    mx = torch.empty_like(avg)
    _call("smaat_cbam_pool_fwd", 4 * B * Cc * H * W, 0, _lib.load().smaat_cbam_pool_fwd, _ptr(x), _ptr(avg), _ptr(mx), B * Cc, H * W, _stream())
    return avg, mx


def cbam_pool_maxpool(x):
    """(avg, mx, maxpool2(x)) from one read of x, or None when the shape is not taken (odd H, W % 4 != 0)."""
    x = _dense(x, "x")
    B, Cc, H, W = x.shape
    if W % 4 != 0 or H % 2 != 0:
        return None
    avg = torch.empty((B, Cc), device=x.device, dtype=torch.float32)
    mx = torch.empty_like(avg)
    pooled = torch.empty((B, Cc, H // 2, W // 2), device=x.device, dtype=torch.float32)
    _call("smaat_cbam_pool_maxpool_fwd", 5 * B * Cc * H * W, 0, _lib.load().smaat_cbam_pool_maxpool_fwd, _ptr(x), _ptr(avg), _ptr(mx),
          _ptr(pooled), B * Cc, H, W, _stream())
    return avg, mx, pooled


_cbam_counters = {}


def _counters(device, n):
    """Zeroed int32 scratch (>= n entries) for the last-arriving-CTA hand-off of smaat_cbam_pool_mlp_fwd: the kernel returns
    it at zero, so one buffer per device serves every call (stream-ordered; allocated outside any graph capture)."""
    t = _cbam_counters.get(device)
    if t is None or t.numel() < n:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("smaat_unet_b200: run one eager forward before capturing a CUDA graph (CBAM scratch allocation)")
        t = torch.zeros(max(n, 256), device=device, dtype=torch.int32)
        _cbam_counters[device] = t
    return t


def cbam_pool_mlp(x, w1, b1, w2, b2, with_maxpool=False):
    """ChannelAttention gate in ONE launch: (sc, avg, mx, pooled or None); None when the shape is not taken
    (C % 8, C > 512, hidden > 64) -- callers then use cbam_pool / cbam_pool_maxpool + cbam_mlp."""
    x = _dense(x, "x")
    B, Cc, H, W = x.shape
    hidden = w1.shape[0]
    if Cc % 8 != 0 or Cc > 512 or hidden > 64:
        return None
    want_pool = with_maxpool and W % 4 == 0 and H % 2 == 0
    avg = torch.empty((B, Cc), device=x.device, dtype=torch.float32)
    mx = torch.empty_like(avg)
    sc = torch.empty_like(avg)
    pooled = torch.empty((B, Cc, H // 2, W // 2), device=x.device, dtype=torch.float32) if want_pool else None
    cnt = _counters(x.device, B)
    _call(f"smaat_cbam_pool_mlp_fwd[C{Cc}_S{H}]", (5 if want_pool else 4) * B * Cc * H * W, 0, _lib.load().smaat_cbam_pool_mlp_fwd, _ptr(x), _ptr(avg), _ptr(mx),
          _ptr(pooled), _ptr(_dense(w1, "w1")), _ptr(b1), _ptr(_dense(w2, "w2")), _ptr(b2), _ptr(sc), _ptr(cnt), B, Cc, H, W, hidden, _stream())
    return sc, avg, mx, pooled


def cbam_gate_scale(x, sc, pooled, wsp, bn_affine, out=None):
    """y = (x * sc) * sigmoid(bn(conv(pooled))) in one launch (layers.py:126-128, :110); None when the shape is not taken."""
    x = _dense(x, "x")
    B, Cc, H, W = x.shape
    if W % 4 != 0:
        return None
    if out is None:
        out = torch.empty_like(x)
        ybs = Cc * H * W
    else:
        out, ybs = _nchw_bstride(out, "out")
    if ybs % 4 != 0 or out.data_ptr() % 16 or x.data_ptr() % 16:
        return None
    ks = wsp.shape[-1]
    _call(f"smaat_cbam_gate_scale_fwd[C{Cc}_S{H}]", 4 * B * (2 * Cc + 2) * H * W, 0, _lib.load().smaat_cbam_gate_scale_fwd, _ptr(pooled), _ptr(_dense(wsp, "wsp")),
          _ptr(bn_affine), _ptr(x), _ptr(sc), _ptr(out), ybs, B, Cc, H, W, ks, _stream())
    return out


def cbam_mlp(avg, mx, w1, b1, w2, b2):
    B, Cc = avg.shape
    sc = torch.empty_like(avg)
    _call("smaat_cbam_mlp_fwd", 12 * B * Cc, 0, _lib.load().smaat_cbam_mlp_fwd, _ptr(avg), _ptr(mx), _ptr(_dense(w1, "w1")), _ptr(b1), _ptr(_dense(w2, "w2")), _ptr(b2),
                                              _ptr(sc), B, Cc, w1.shape[0], _stream())
    return sc


def cbam_reduce(x, sc):
    x = _dense(x, "x")
    B, Cc, H, W = x.shape
    pooled = torch.empty((B, 2, H, W), device=x.device, dtype=torch.float32)
    _call(f"smaat_cbam_reduce_fwd[C{Cc}_S{H}]", 4 * B * (Cc + 2) * H * W, 0, _lib.load().smaat_cbam_reduce_fwd, _ptr(x), _ptr(sc), _ptr(pooled), B, Cc, H * W, _stream())
    return pooled


def cbam_gate(pooled, wsp, bn_affine, want_raw=False):
    B, _, H, W = pooled.shape
    sa = torch.empty((B, 1, H, W), device=pooled.device, dtype=torch.float32)
    raw = torch.empty_like(sa) if want_raw else None
    ks = wsp.shape[-1]
    _call("smaat_cbam_gate_fwd", 12 * B * H * W, 0, _lib.load().smaat_cbam_gate_fwd, _ptr(pooled), _ptr(_dense(wsp, "wsp")), _ptr(bn_affine), _ptr(sa), _ptr(raw), B, H, W, ks,
                                               _stream())
    return (sa, raw) if want_raw else sa


def cbam_scale(x, sc, sa, out=None):
    x = _dense(x, "x")
    B, Cc, H, W = x.shape
    if out is None:
        out = torch.empty_like(x)
        ybs = Cc * H * W
    else:
        out, ybs = _nchw_bstride(out, "out")
    _call("smaat_cbam_scale_fwd", 4 * B * (2 * Cc + 1) * H * W, 0, _lib.load().smaat_cbam_scale_fwd, _ptr(x), _ptr(sc), _ptr(sa), _ptr(out), ybs, B, Cc, H * W, _stream())
    return out


def outconv(x, weight, bias):
    """OutConv 1x1 (unet_parts.py:70)."""
    x = _dense(x, "x")
    B, Cin, H, W = x.shape
    w = _dense(weight, "weight")
    ncls = w.shape[0]
    y = torch.empty((B, ncls, H, W), device=x.device, dtype=torch.float32)
    _call("smaat_outconv_fwd", 4 * B * (Cin + ncls) * H * W, 2 * B * Cin * ncls * H * W, _lib.load().smaat_outconv_fwd, _ptr(x), _ptr(w), _ptr(bias), _ptr(y), B, Cin, ncls, H * W, _stream())
    return y
