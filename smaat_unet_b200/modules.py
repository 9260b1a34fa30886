"""Drop-in nn.Module replacements for the SmaAt-UNet hot-path blocks.

Host-side mirror of the reference's class interface: same class names,
constructor signatures, forward signatures and ``state_dict`` keys as

* ``models/layers.py``: DepthwiseSeparableConv (:34-50), ChannelAttention (:90-111),
  SpatialAttention (:114-129), CBAM (:132-141)
* ``models/unet_parts_depthwise_separable.py``: DoubleConvDS (:10-39), DownDS (:42-53), UpDS (:56-86)
* ``models/unet_parts.py``: OutConv (:67-73)

so checkpoints trained with the reference load unchanged (``calc_metrics_test_set.py:114``).
The parameter containers are ordinary torch modules (``nn.Conv2d`` / ``nn.BatchNorm2d`` /
``nn.Linear`` -> identical default initialisation and key names) but their ``forward`` is
never called: all arithmetic runs in libsmaat_b200.so (sm_100a kernels) through ``ops``.
There is no PyTorch / CPU fallback; unsupported requests raise.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


def _versions(*tensors):
    """Cache key for tensors derived from parameters / buffers.  ``_version`` only sees writes that go through torch
    dispatch; our kernels (BatchNorm running statistics in smaat_bn_finalize, anything replayed from a CUDA graph)
    write through raw pointers, so the key also carries ``ops.weights_generation()`` -- bumped by every such write
    (ops.bn_finalize, TrainSession.step) and by every train()/eval() switch of a caching module."""
    return (ops.weights_generation(),) + tuple((t.data_ptr(), t._version) for t in tensors if t is not None)


class _CachingModule(nn.Module):
    """Modules that cache tensors derived from their parameters drop those caches on every mode switch."""

    def _drop_caches(self):
        pass

    def train(self, mode=True):
        self._drop_caches()
        ops.bump_weights_generation()
        return super().train(mode)


def _needs_grad(mod, *inputs):
    """True when the call must be recorded on the autograd tape (grad mode on and a parameter or input requires grad)."""
    if not torch.is_grad_enabled():
        return False
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in inputs):
        return True
    return any(p.requires_grad for p in mod.parameters())


def _no_autograd(mod, *inputs):
    if _needs_grad(mod, *inputs):
        raise NotImplementedError(
            f"{type(mod).__name__}: autograd through this standalone module is not implemented in smaat_unet_b200 "
            "(DoubleConvDS / DownDS / UpDS / CBAM / OutConv are differentiable). Use torch.no_grad(); no PyTorch fallback is provided.")


class DepthwiseSeparableConv(_CachingModule):
    """models/layers.py:34-50 -- depthwise(k x k, groups=Cin, kpl outputs per channel) then pointwise 1x1."""

    def __init__(self, in_channels, output_channels, kernel_size, padding=0, kernels_per_layer=1):
        super().__init__()
        self.depthwise = nn.Conv2d(in_channels, in_channels * kernels_per_layer, kernel_size=kernel_size,
                                   padding=padding, groups=in_channels)
        self.pointwise = nn.Conv2d(in_channels * kernels_per_layer, output_channels, kernel_size=1)
        self.kernels_per_layer = kernels_per_layer
        self._wsplit = None
        self._wsplit_key = None

    def _drop_caches(self):
        self._wsplit = self._wsplit_key = None

    def _check(self):
        if self.depthwise.kernel_size != (3, 3) or self.depthwise.padding != (1, 1):
            raise NotImplementedError("smaat_unet_b200 implements the depthwise conv the reference uses: 3x3, padding=1 "
                                      "(parts_ds.py:18-33)")

    def pw_split(self):
        """(hi, lo) tf32 split of the pointwise weight, cached on the parameter's version counter."""
        w = self.pointwise.weight
        key = _versions(w)
        # while a training step is being captured into a CUDA graph the split must be part of the graph: replays see new weights
        in_train_capture = self.training and torch.cuda.is_current_stream_capturing()
        if in_train_capture or self._wsplit_key != key:
            with torch.no_grad():
                self._wsplit = ops.split_tf32(w.detach().view(w.shape[0], -1))
            self._wsplit_key = None if in_train_capture else key
        return self._wsplit

    def fused_takes(self, x, x1=None, stats=False) -> bool:
        """Whether the one-kernel depthwise->pointwise path takes this input (shape, alignment, arithmetic mode)."""
        return ops.dsconv_takes(x, x1, self.pointwise.weight.detach(), self.kernels_per_layer, stats=stats)

    def run(self, x, x1=None, scale=None, shift=None, relu=False, in_scale=None, in_shift=None, stats=None, outconv=None):
        """dw -> pw with the pw epilogue y = act(scale * acc + shift).  scale/shift None => (1, pointwise.bias)."""
        self._check()
        dw_b = self.depthwise.bias.detach() if self.depthwise.bias is not None else None
        if shift is None:
            shift = self.pointwise.bias.detach() if self.pointwise.bias is not None else None
        mode = ops.get_pointwise_mode()
        split = self.pw_split() if mode == "tf32x3" else None
        if in_scale is None:
            # one kernel: the k*Cin-channel depthwise result never reaches HBM (where the shape allows)
            y = ops.dsconv(x, self.depthwise.weight.detach(), dw_b, self.kernels_per_layer, self.pointwise.weight.detach(),
                           scale, shift, relu, x1=x1, mode=mode, w_split=split, stats=stats, outconv=outconv)
            if y is not None:
                return y
        if outconv is not None:
            return None      # the caller runs this conv and the OutConv separately
        d = ops.dw3x3(x, self.depthwise.weight.detach(), dw_b, self.kernels_per_layer, x1=x1, in_scale=in_scale, in_shift=in_shift)
        return ops.pw1x1(d, self.pointwise.weight.detach(), scale, shift, relu, mode=mode, w_split=split, stats=stats)

    def forward(self, x):
        if _needs_grad(self, x):
            from .autograd import DSConvFn
            self._check()
            return DSConvFn.run(self, x)
        return self.run(x)


class DoubleConvDS(_CachingModule):
    """models/unet_parts_depthwise_separable.py:10-39 -- (DS conv => BN => ReLU) * 2.

    Eval mode runs 4 kernels: dw, pw(+folded BN +ReLU), dw, pw(+folded BN +ReLU).
    """

    def __init__(self, in_channels, out_channels, mid_channels=None, kernels_per_layer=1):
        super().__init__()
        if not mid_channels:
            mid_channels = out_channels
        self.double_conv = nn.Sequential(
            DepthwiseSeparableConv(in_channels, mid_channels, kernel_size=3, kernels_per_layer=kernels_per_layer, padding=1),
            nn.BatchNorm2d(mid_channels),
            nn.ReLU(inplace=True),
            DepthwiseSeparableConv(mid_channels, out_channels, kernel_size=3, kernels_per_layer=kernels_per_layer, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
        )
        self._fold = {}

    def _drop_caches(self):
        self._fold = {}

    def _folded(self, idx):
        """(scale, shift) of eval BatchNorm idx+1 folded with pointwise bias of DS conv idx; cached."""
        ds, bn = self.double_conv[idx], self.double_conv[idx + 1]
        pb = ds.pointwise.bias
        key = _versions(bn.weight, bn.bias, bn.running_mean, bn.running_var, pb)
        hit = self._fold.get(idx)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                sc_sh = ops.bn_fold(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                                    pb.detach() if pb is not None else None, bn.eps)
            self._fold[idx] = (key, sc_sh)
            hit = self._fold[idx]
        return hit[1]

    def run(self, x, x1=None, outconv=None):
        """``outconv`` (an OutConv module with one class, inference only): fold it into the last kernel's epilogue and
        return the logits -- the block's own output is then never materialised (models/SmaAt_UNet.py:55-56)."""
        ops._req(x, "input", 4)
        if outconv is not None:
            y = self._run_with_outconv(x, x1, outconv)
            return y if y is not None else outconv(self.run(x, x1))
        if _needs_grad(self, x, x1):
            from .autograd import DoubleConvDSFn
            return DoubleConvDSFn.run(self, x, x1)
        bns = (self.double_conv[1], self.double_conv[4])
        if self.training or any(not bn.track_running_stats or bn.running_mean is None for bn in bns):
            from . import functional as Fn       # batch statistics (and running-stat update), no tape
            return Fn.double_conv_fwd(self, x, x1)[0]
        s0, t0 = self._folded(0)
        y = self.double_conv[0].run(x, x1=x1, scale=s0, shift=t0, relu=True)
        s1, t1 = self._folded(3)
        return self.double_conv[3].run(y, scale=s1, shift=t1, relu=True)

    def _run_with_outconv(self, x, x1, outconv):
        bns = (self.double_conv[1], self.double_conv[4])
        oc = outconv.conv
        if (_needs_grad(self, x, x1) or _needs_grad(outconv, x) or self.training or oc.out_channels != 1
                or any(not bn.track_running_stats or bn.running_mean is None for bn in bns)):
            return None
        s0, t0 = self._folded(0)
        y = self.double_conv[0].run(x, x1=x1, scale=s0, shift=t0, relu=True)
        s1, t1 = self._folded(3)
        ob = oc.bias.detach() if oc.bias is not None else None
        z = self.double_conv[3].run(y, scale=s1, shift=t1, relu=True, outconv=(oc.weight.detach(), ob))
        if z is None:     # shape / mode not taken by the fused kernel: same two convs, OutConv as its own kernel
            z = outconv(self.double_conv[3].run(y, scale=s1, shift=t1, relu=True))
        return z

    def forward(self, x):
        return self.run(x)


# The reference calls ``cbamN(x)`` and then ``downN(x)`` on the same un-attended map (models/SmaAt_UNet.py:42-50,
# unet_precip_regression_lightning.py:148-157): the channel gate's global pools and the 2x2 max-pool are produced by ONE
# read of x.  The plain-call API has no way to hand the pooled map over, so the CBAM leaves it in this one-slot stash,
# keyed on the identity of x; DownDS takes it when (and only when) it is called on the very same tensor.
_maxpool_stash = None


def _tensor_key(x):
    return (x.data_ptr(), x._version, tuple(x.shape), tuple(x.stride()), x.device.index)


def _stash_maxpool(x, pooled):
    global _maxpool_stash
    _maxpool_stash = (_tensor_key(x), pooled)


def _take_stashed_maxpool(x):
    global _maxpool_stash
    hit = _maxpool_stash
    if hit is None or not isinstance(x, torch.Tensor) or not x.is_cuda:
        return None
    if hit[0] != _tensor_key(x):
        return None
    _maxpool_stash = None
    return hit[1]


class DownDS(nn.Module):
    """models/unet_parts_depthwise_separable.py:42-53 -- MaxPool2d(2) then DoubleConvDS."""

    def __init__(self, in_channels, out_channels, kernels_per_layer=1):
        super().__init__()
        self.maxpool_conv = nn.Sequential(
            nn.MaxPool2d(2),
            DoubleConvDS(in_channels, out_channels, kernels_per_layer=kernels_per_layer),
        )

    def forward(self, x, pooled=None):
        """``pooled``: MaxPool2d(2)(x) when the caller already has it.  Called plainly (``down(x)``, as the reference does)
        it first looks for the 2x2 max-pool the preceding ``CBAM(x)`` call left behind (see ``CBAM.forward``)."""
        if pooled is None:
            pooled = _take_stashed_maxpool(x)
        if pooled is None:
            if _needs_grad(self, x):
                from .autograd import MaxPool2Fn
                pooled = MaxPool2Fn.apply(x) if x.requires_grad else ops.maxpool2(x)
            else:
                pooled = ops.maxpool2(x)
        return self.maxpool_conv[1].run(pooled)


class UpDS(_CachingModule):
    """models/unet_parts_depthwise_separable.py:56-86 -- upsample x2, pad to the skip, concat, DoubleConvDS.

    The concat is never materialised: the first depthwise kernel reads [skip, up] as a virtual concat.
    ``bilinear=False`` (ConvTranspose2d(in, in // 2, 2, stride=2), :72-73): kernel = stride, so the transposed conv is one
    tcgen05 pointwise GEMM to the 4 packed taps + a pixel shuffle (csrc/convt.cu).
    """

    def __init__(self, in_channels, out_channels, bilinear=True, kernels_per_layer=1):
        super().__init__()
        self.bilinear = bilinear
        if bilinear:
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
            self.conv = DoubleConvDS(in_channels, out_channels, in_channels // 2, kernels_per_layer=kernels_per_layer)
        else:
            self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
            self.conv = DoubleConvDS(in_channels, out_channels, kernels_per_layer=kernels_per_layer)
        self._packed = None

    def _drop_caches(self):
        self._packed = None

    def _packed_weight(self):
        """((4 Cout, Cin) GEMM matrix of the transposed conv, its tf32 (hi, lo) split or None); cached on the weight's version."""
        w = self.up.weight
        key = (_versions(w), ops.get_pointwise_mode())
        in_train_capture = self.training and torch.cuda.is_current_stream_capturing()
        if in_train_capture or self._packed is None or self._packed[0] != key:
            with torch.no_grad():
                wp = ops.convt2x2_pack_weight(w.detach())
                split = ops.split_tf32(wp) if ops.get_pointwise_mode() == "tf32x3" else None
            self._packed = (None if in_train_capture else key, wp, split)
        return self._packed[1], self._packed[2]

    def _up_transposed(self, x1, Ho, Wo):
        up = self.up
        if (up.kernel_size, up.stride, up.padding, up.output_padding, up.dilation, up.groups) != ((2, 2), (2, 2), (0, 0), (0, 0), (1, 1), 1):
            raise NotImplementedError("UpDS: only the reference's ConvTranspose2d(kernel_size=2, stride=2) is implemented (parts_ds.py:72)")
        wp, split = self._packed_weight()
        if _needs_grad(up, x1):
            from .autograd import ConvT2x2PadFn
            return ConvT2x2PadFn.apply(x1, up.weight, up.bias, Ho, Wo, wp, split)
        t = ops.pw1x1(x1, wp, None, None, False, w_split=split)
        return ops.pixel_shuffle2_pad(t, up.bias.detach() if up.bias is not None else None, up.out_channels, Ho, Wo)

    def forward(self, x1, x2, outconv=None):
        if not self.bilinear:
            return self.conv.run(x2, x1=self._up_transposed(x1, x2.shape[2], x2.shape[3]), outconv=outconv)
        if torch.is_grad_enabled() and x1.requires_grad:
            from .autograd import Upsample2xPadFn
            up = Upsample2xPadFn.apply(x1, x2.shape[2], x2.shape[3])
        else:
            up = ops.upsample2x_pad(x1, x2.shape[2], x2.shape[3])
        return self.conv.run(x2, x1=up, outconv=outconv)


class OutConv(nn.Module):
    """models/unet_parts.py:67-73 -- 1x1 conv to n_classes."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)

    def forward(self, x):
        if _needs_grad(self, x):
            from .autograd import OutConvFn
            return OutConvFn.apply(x, self.conv.weight, self.conv.bias)
        return ops.outconv(x, self.conv.weight.detach(), self.conv.bias.detach() if self.conv.bias is not None else None)


class Flatten(nn.Module):
    """models/layers.py:85-87 (kept so that MLP indices -- state_dict keys MLP.1 / MLP.3 -- match)."""

    def forward(self, x):
        return x.view(x.size(0), -1)


class ChannelAttention(nn.Module):
    """models/layers.py:90-111."""

    def __init__(self, input_channels, reduction_ratio=16):
        super().__init__()
        self.input_channels = input_channels
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.max_pool = nn.AdaptiveMaxPool2d(1)
        self.MLP = nn.Sequential(
            Flatten(),
            nn.Linear(input_channels, input_channels // reduction_ratio),
            nn.ReLU(),
            nn.Linear(input_channels // reduction_ratio, input_channels),
        )

    def gate(self, x, with_maxpool=False):
        """sigmoid(MLP(avg) + MLP(max)) as a (B, C) tensor.  ``with_maxpool``: also return MaxPool2d(2)(x) (or None), computed
        in the same read of x as the global pools."""
        l1, l2 = self.MLP[1], self.MLP[3]
        # 512 channels: the MLP run by the last-arriving pooling CTA of each image measured SLOWER than a second launch
        # (tools/time_cbam_pool.py, B = 32: 57.6 vs 24.5 us at 18 x 18, 49.8 vs 37.7 us at 36 x 36; equal from 72 x 72 up)
        one = None if x.shape[1] >= 512 else ops.cbam_pool_mlp(x, l1.weight.detach(), l1.bias.detach(), l2.weight.detach(), l2.bias.detach(),
                                                               with_maxpool=with_maxpool)
        if one is not None:          # pools + MLP + sigmoid (+ the 2x2 max-pool) in one launch
            sc, _, _, pooled = one
            return (sc, pooled) if with_maxpool else sc
        pooled = None
        fused = ops.cbam_pool_maxpool(x) if with_maxpool else None
        if fused is not None:
            avg, mx, pooled = fused
        else:
            avg, mx = ops.cbam_pool(x)
        sc = ops.cbam_mlp(avg, mx, l1.weight.detach(), l1.bias.detach(), l2.weight.detach(), l2.bias.detach())
        return (sc, pooled) if with_maxpool else sc

    def forward(self, x):
        _no_autograd(self, x)
        sc = self.gate(x)
        ones = torch.ones((x.shape[0], 1, x.shape[2], x.shape[3]), device=x.device, dtype=torch.float32)
        return ops.cbam_scale(x, sc, ones)


class SpatialAttention(_CachingModule):
    """models/layers.py:114-129."""

    def __init__(self, kernel_size=7):
        super().__init__()
        assert kernel_size in (3, 7), "kernel size must be 3 or 7"
        padding = 3 if kernel_size == 7 else 1
        self.conv = nn.Conv2d(2, 1, kernel_size=kernel_size, padding=padding, bias=False)
        self.bn = nn.BatchNorm2d(1)
        self._fold = None

    def _drop_caches(self):
        self._fold = None

    def bn_affine(self):
        """Device tensor [scale, shift] of the eval-mode BatchNorm2d(1); cached."""
        if self.bn.training:
            raise NotImplementedError("standalone SpatialAttention in train mode is not implemented (CBAM handles it); "
                                      "call .eval() or use CBAM")
        bn = self.bn
        key = _versions(bn.weight, bn.bias, bn.running_mean, bn.running_var)
        if self._fold is None or self._fold[0] != key:
            with torch.no_grad():
                s, t = ops.bn_fold(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, None, bn.eps)
                self._fold = (key, torch.cat([s, t]))
        return self._fold[1]

    def gate(self, x, sc):
        pooled = ops.cbam_reduce(x, sc)
        return ops.cbam_gate(pooled, self.conv.weight.detach(), self.bn_affine())

    def forward(self, x):
        _no_autograd(self, x)
        ones = torch.ones((x.shape[0], x.shape[1]), device=x.device, dtype=torch.float32)
        return ops.cbam_scale(x, ones, self.gate(x, ones))


class CBAM(nn.Module):
    """models/layers.py:132-141 -- channel attention then spatial attention: 3 kernels, 4|x| of traffic."""

    def __init__(self, input_channels, reduction_ratio=16, kernel_size=7):
        super().__init__()
        self.channel_att = ChannelAttention(input_channels, reduction_ratio=reduction_ratio)
        self.spatial_att = SpatialAttention(kernel_size=kernel_size)
        self.stash_maxpool = True     # False: never produce the max-pool in a plain call (e.g. no DownDS follows)

    def forward(self, x, out=None, with_maxpool=False):
        """``with_maxpool=True`` returns (CBAM(x), MaxPool2d(2)(x) or None) explicitly; a plain ``cbam(x)`` (the reference's
        call) returns CBAM(x) and stashes the max-pool for the DownDS that follows (models/SmaAt_UNet.py:42-50)."""
        if _needs_grad(self, x):
            from .autograd import CBAMFn
            y = CBAMFn.run(self, x)
            return (y, None) if with_maxpool else y
        if self.spatial_att.bn.training or not self.spatial_att.bn.track_running_stats:
            from . import functional as Fn       # batch statistics for the gate's BatchNorm2d(1), no tape
            y = Fn.cbam_fwd(self, ops._dense(x, "x"))[0]
            return (y, None) if with_maxpool else y
        # one read of x gives the global pools AND MaxPool2d(2)(x) (shape permitting); a plain call leaves the latter for
        # the DownDS that the reference calls next on the same x
        pooled = None
        if with_maxpool or self.stash_maxpool:
            sc, pooled = self.channel_att.gate(x, with_maxpool=True)
            if pooled is not None and not with_maxpool:
                _stash_maxpool(x, pooled)
        else:
            sc = self.channel_att.gate(x)
        # three launches: [pools + MLP (+ max-pool)], channel reduce, [k x k gate + scale]
        red = ops.cbam_reduce(x, sc)
        y = ops.cbam_gate_scale(x, sc, red, self.spatial_att.conv.weight.detach(), self.spatial_att.bn_affine(), out=out)
        if y is None:
            sa = ops.cbam_gate(red, self.spatial_att.conv.weight.detach(), self.spatial_att.bn_affine())
            y = ops.cbam_scale(x, sc, sa, out=out)
        return (y, pooled) if with_maxpool else y


def cached_tensors(model):
    """Every tensor the eval fast path derived from ``model``'s parameters (folded BatchNorm affine, tf32 hi/lo splits).
    A CUDA graph captured over the eval forward has their addresses baked in: whoever owns the graph must keep them alive."""
    out = []
    for m in model.modules():
        if isinstance(m, DepthwiseSeparableConv) and m._wsplit is not None:
            out.extend(m._wsplit)
        elif isinstance(m, DoubleConvDS):
            for _, pair in m._fold.values():
                out.extend(pair)
        elif isinstance(m, SpatialAttention) and m._fold is not None:
            out.append(m._fold[1])
    return out
