"""CPU oracle for the SmaAt-UNet forward hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-numpy restatement of the reference's algorithm for the
path named in BASELINE.json (DS-conv blocks + CBAM + the glue between them).
It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product path (``smaat_unet_b200``) never does, and has no CPU fallback.

Pinning: the reference ships no tests and no golden vectors ("parity
unpinned" in SURVEY.md section 4/8c), so the oracle is pinned against outputs
of the reference itself, executed in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/``; see
``tests/test_oracle_golden.py``.  Every function cites the reference lines
(relative to /root/reference) that it restates.

All functions are dtype-generic: feed float64 arrays for a tight algorithmic
check, float32 arrays for "what an fp32 implementation should produce".
Arrays are NCHW, C-contiguous.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-5        # torch.nn.BatchNorm2d default, used at parts_ds.py:25,34 and layers.py:120
BN_MOMENTUM = 0.1    # torch.nn.BatchNorm2d default


# ----------------------------------------------------------------------------
# Leaf arithmetic
# ----------------------------------------------------------------------------
def depthwise3x3(x, weight, bias, kernels_per_layer):
    """models/layers.py:38-44,48 -- Conv2d(Cin, k*Cin, 3, padding=1, groups=Cin).

    Output channel ``o`` reads input channel ``o // k`` (grouped conv with
    Cin groups), cross-correlation, zero padding 1, plus bias[o].
    weight: (k*Cin, 1, 3, 3)   bias: (k*Cin,) or None
    """
    B, C, H, W = x.shape
    k = int(kernels_per_layer)
    assert weight.shape == (k * C, 1, 3, 3), weight.shape
    xp = np.zeros((B, C, H + 2, W + 2), dtype=x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    src = np.repeat(xp, k, axis=1) if k > 1 else xp            # out channel o <- in channel o//k
    y = np.zeros((B, k * C, H, W), dtype=x.dtype)
    w = weight.reshape(k * C, 3, 3).astype(x.dtype)
    for dy in range(3):
        for dx in range(3):
            y += w[None, :, dy, dx, None, None] * src[:, :, dy:dy + H, dx:dx + W]
    if bias is not None:
        y += bias.astype(x.dtype)[None, :, None, None]
    return y


def pointwise1x1(x, weight, bias):
    """models/layers.py:45,49 and models/unet_parts.py:70 -- Conv2d(K, Cout, 1).

    y[b,o,p] = bias[o] + sum_c W[o,c] x[b,c,p];  weight: (Cout, K, 1, 1).
    """
    B, K, H, W = x.shape
    Wm = weight.reshape(weight.shape[0], K).astype(x.dtype)
    y = np.matmul(Wm[None], x.reshape(B, K, H * W)).reshape(B, Wm.shape[0], H, W)
    if bias is not None:
        y = y + bias.astype(x.dtype)[None, :, None, None]
    return y


def batchnorm_eval(x, gamma, beta, running_mean, running_var, eps=BN_EPS):
    """nn.BatchNorm2d in eval mode (parts_ds.py:25,34; layers.py:120,127)."""
    d = x.dtype
    inv = 1.0 / np.sqrt(running_var.astype(d) + d.type(eps))
    return (x - running_mean.astype(d)[None, :, None, None]) * (inv * gamma.astype(d))[None, :, None, None] \
        + beta.astype(d)[None, :, None, None]


def batchnorm_train(x, gamma, beta, running_mean, running_var, eps=BN_EPS, momentum=BN_MOMENTUM):
    """nn.BatchNorm2d in train mode: batch mean / *biased* variance over (B,H,W)
    normalise; running stats get the *unbiased* variance.

    Returns (y, new_running_mean, new_running_var).
    """
    d = x.dtype
    n = x.shape[0] * x.shape[2] * x.shape[3]
    mean = x.mean(axis=(0, 2, 3))
    var = x.var(axis=(0, 2, 3))                                 # biased
    inv = 1.0 / np.sqrt(var + d.type(eps))
    y = (x - mean[None, :, None, None]) * (inv * gamma.astype(d))[None, :, None, None] \
        + beta.astype(d)[None, :, None, None]
    unbiased = var * (n / max(n - 1, 1))
    new_rm = (1 - momentum) * running_mean.astype(d) + momentum * mean
    new_rv = (1 - momentum) * running_var.astype(d) + momentum * unbiased
    return y, new_rm, new_rv


def relu(x):
    """nn.ReLU (parts_ds.py:26,35; layers.py:100)."""
    return np.maximum(x, 0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def maxpool2(x):
    """nn.MaxPool2d(2) (parts_ds.py:48): stride 2, floor mode."""
    B, C, H, W = x.shape
    Ho, Wo = H // 2, W // 2
    v = x[:, :, :2 * Ho, :2 * Wo].reshape(B, C, Ho, 2, Wo, 2)
    return v.max(axis=(3, 5))


def upsample_bilinear2x(x):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
    (parts_ds.py:64): src = dst * (in-1)/(out-1), linear interpolation.
    """
    B, C, H, W = x.shape
    Ho, Wo = 2 * H, 2 * W
    d = x.dtype

    def axis_weights(n_in, n_out):
        scale = d.type(n_in - 1) / d.type(n_out - 1) if n_out > 1 else d.type(0)
        src = np.arange(n_out, dtype=d) * scale
        i0 = np.floor(src).astype(np.int64)
        i0 = np.minimum(i0, n_in - 1)
        i1 = np.minimum(i0 + 1, n_in - 1)
        lam = (src - i0.astype(d)).astype(d)
        return i0, i1, lam

    y0, y1, ly = axis_weights(H, Ho)
    x0, x1, lx = axis_weights(W, Wo)
    top = x[:, :, y0, :]
    bot = x[:, :, y1, :]
    rows = top * (1 - ly)[None, None, :, None] + bot * ly[None, None, :, None]
    left = rows[:, :, :, x0]
    right = rows[:, :, :, x1]
    return left * (1 - lx)[None, None, None, :] + right * lx[None, None, None, :]


def pad_to(x, H, W):
    """F.pad(x1, [dX//2, dX-dX//2, dY//2, dY-dY//2]) (parts_ds.py:78-81), zero fill."""
    dY = H - x.shape[2]
    dX = W - x.shape[3]
    assert dY >= 0 and dX >= 0, "reference only pads (negative pad would crop)"
    return np.pad(x, ((0, 0), (0, 0), (dY // 2, dY - dY // 2), (dX // 2, dX - dX // 2)))


def conv2d_same(x, weight, pad):
    """Small dense conv for SpatialAttention (layers.py:119): (Cout,Cin,k,k), no bias."""
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    y = np.zeros((B, Cout, H, W), dtype=x.dtype)
    w = weight.astype(x.dtype)
    for o in range(Cout):
        for c in range(Cin):
            for dy in range(kh):
                for dx in range(kw):
                    y[:, o] += w[o, c, dy, dx] * xp[:, c, dy:dy + H, dx:dx + W]
    return y


# ----------------------------------------------------------------------------
# State-dict helpers (keys are the reference's)
# ----------------------------------------------------------------------------
def _sub(sd, prefix):
    p = prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


def _bn(x, sd, prefix, training):
    g, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        y, nrm, nrv = batchnorm_train(x, g, b, rm, rv)
        return y, {prefix + ".running_mean": nrm, prefix + ".running_var": nrv,
                   prefix + ".num_batches_tracked": np.asarray(sd.get(prefix + ".num_batches_tracked", 0)) + 1}
    return batchnorm_eval(x, g, b, rm, rv), {}


# ----------------------------------------------------------------------------
# Blocks
# ----------------------------------------------------------------------------
def ds_conv(x, sd, prefix, k):
    """DepthwiseSeparableConv.forward (layers.py:47-50): depthwise then pointwise, nothing between."""
    y = depthwise3x3(x, sd[prefix + ".depthwise.weight"], sd[prefix + ".depthwise.bias"], k)
    return pointwise1x1(y, sd[prefix + ".pointwise.weight"], sd[prefix + ".pointwise.bias"])


def double_conv_ds(x, sd, prefix, k, training=False):
    """DoubleConvDS (parts_ds.py:10-39): [DS, BN, ReLU] x 2 as double_conv.{0,1,2,3,4,5}."""
    upd = {}
    y = ds_conv(x, sd, prefix + ".double_conv.0", k)
    y, u = _bn(y, sd, prefix + ".double_conv.1", training); upd.update(u)
    y = relu(y)
    y = ds_conv(y, sd, prefix + ".double_conv.3", k)
    y, u = _bn(y, sd, prefix + ".double_conv.4", training); upd.update(u)
    return relu(y), upd


def down_ds(x, sd, prefix, k, training=False):
    """DownDS (parts_ds.py:42-53): MaxPool2d(2) then DoubleConvDS under maxpool_conv.1."""
    return double_conv_ds(maxpool2(x), sd, prefix + ".maxpool_conv.1", k, training)


def conv_transpose2x2(x, w, b):
    """nn.ConvTranspose2d(Cin, Cout, kernel_size=2, stride=2) (parts_ds.py:72): w (Cin, Cout, 2, 2);
    y[b, o, 2i + dy, 2j + dx] = bias[o] + sum_c x[b, c, i, j] * w[c, o, dy, dx] -- no overlapping taps."""
    B, _, H, W = x.shape
    Cout = w.shape[1]
    y = np.zeros((B, Cout, 2 * H, 2 * W), dtype=x.dtype)
    for dy in range(2):
        for dx in range(2):
            y[:, :, dy::2, dx::2] = np.einsum("bcij,co->boij", x, w[:, :, dy, dx].astype(x.dtype))
    return y + b.astype(x.dtype)[None, :, None, None]


def up_ds(x_low, x_skip, sd, prefix, k, training=False):
    """UpDS (parts_ds.py:56-86): upsample x2 (bilinear, or ConvTranspose2d when the state_dict holds `up.weight`), pad to skip,
    cat([skip, up]), DoubleConvDS."""
    if prefix + ".up.weight" in sd:
        up = conv_transpose2x2(x_low, sd[prefix + ".up.weight"], sd[prefix + ".up.bias"])
    else:
        up = upsample_bilinear2x(x_low)
    up = pad_to(up, x_skip.shape[2], x_skip.shape[3])
    return double_conv_ds(np.concatenate([x_skip, up], axis=1), sd, prefix + ".conv", k, training)


def channel_attention(x, sd, prefix):
    """ChannelAttention.forward (layers.py:105-111): shared MLP on avg- and max-pooled
    vectors, summed (second-layer bias therefore counted twice), sigmoid, scale."""
    avg = x.mean(axis=(2, 3))
    mx = x.max(axis=(2, 3))
    W1, b1 = sd[prefix + ".MLP.1.weight"].astype(x.dtype), sd[prefix + ".MLP.1.bias"].astype(x.dtype)
    W2, b2 = sd[prefix + ".MLP.3.weight"].astype(x.dtype), sd[prefix + ".MLP.3.bias"].astype(x.dtype)

    def mlp(v):
        return relu(v @ W1.T + b1) @ W2.T + b2

    s = sigmoid(mlp(avg) + mlp(mx))
    return x * s[:, :, None, None]


def spatial_attention(x, sd, prefix, training=False):
    """SpatialAttention.forward (layers.py:122-129): cat(mean_c, max_c) -> conv kxk (no bias)
    -> BatchNorm2d(1) -> sigmoid -> scale."""
    w = sd[prefix + ".conv.weight"]
    ks = w.shape[-1]
    p = np.concatenate([x.mean(axis=1, keepdims=True), x.max(axis=1, keepdims=True)], axis=1)
    a = conv2d_same(p, w, 3 if ks == 7 else 1)
    a, upd = _bn(a, sd, prefix + ".bn", training)
    return x * sigmoid(a), upd


def cbam(x, sd, prefix, training=False):
    """CBAM.forward (layers.py:138-141): channel attention then spatial attention."""
    return spatial_attention(channel_attention(x, sd, prefix + ".channel_att"), sd, prefix + ".spatial_att", training)


def out_conv(x, sd, prefix):
    """OutConv (unet_parts.py:67-73): 1x1 conv + bias, no activation."""
    return pointwise1x1(x, sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"])


def smaat_unet_forward(x, sd, kernels_per_layer=2, training=False, return_updates=False, n_cbams=5):
    """SmaAt_UNet.forward (models/SmaAt_UNet.py:41-57), bilinear=True.

    Un-attended maps feed the next encoder stage, attended maps are the decoder
    skips, x5Att is the decoder input.  n_cbams = 4: UNetDSAttention4CBAMs (x5 un-attended,
    unet_precip_regression_lightning.py:193-208); n_cbams = 0: UNetDS (:104-117).
    """
    k = kernels_per_layer
    upd = {}

    def acc(res):
        y, u = res
        upd.update(u)
        return y

    enc = [acc(double_conv_ds(x, sd, "inc", k, training))]
    for i in range(1, 5):
        enc.append(acc(down_ds(enc[-1], sd, f"down{i}", k, training)))
    att = [acc(cbam(e, sd, f"cbam{i + 1}", training)) if i < n_cbams else e for i, e in enumerate(enc)]
    y = att[4]
    for i in range(1, 5):
        y = acc(up_ds(y, att[4 - i], sd, f"up{i}", k, training))
    y = out_conv(y, sd, "outc")
    return (y, upd) if return_updates else y
