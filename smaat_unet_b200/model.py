"""SmaAt-UNet assembled from the B200 drop-in blocks.

Same constructor, attribute names (hence state_dict keys) and forward graph as the
reference's ``models/SmaAt_UNet.py:7-57``; provided so the full model can be built where the
reference checkout is not importable (e.g. the GPU box).  With the reference on
``sys.path`` prefer ``smaat_unet_b200.patch_reference()`` and use its own ``SmaAt_UNet``
(and the Lightning wrappers) unchanged.
"""
from __future__ import annotations

from torch import nn

from .modules import CBAM, DoubleConvDS, DownDS, OutConv, UpDS

_ENC = (64, 128, 256, 512)


class SmaAt_UNet(nn.Module):
    def __init__(self, n_channels, n_classes, kernels_per_layer=2, bilinear=True, reduction_ratio=16):
        super().__init__()
        self.n_channels, self.n_classes, self.bilinear = n_channels, n_classes, bilinear
        k, r = kernels_per_layer, reduction_ratio
        factor = 2 if bilinear else 1
        widths = list(_ENC) + [1024 // factor]            # channels of x1..x5
        self.inc = DoubleConvDS(n_channels, widths[0], kernels_per_layer=k)
        for lvl in range(5):                                # [down_l,] cbam_{l+1} -- registration order = reference's
            if lvl > 0:
                setattr(self, f"down{lvl}", DownDS(widths[lvl - 1], widths[lvl], kernels_per_layer=k))
            setattr(self, f"cbam{lvl + 1}", CBAM(widths[lvl], reduction_ratio=r))
        dec_in = (1024, 512, 256, 128)
        dec_out = (512 // factor, 256 // factor, 128 // factor, 64)
        for i in range(4):                                  # up1..up4
            setattr(self, f"up{i + 1}", UpDS(dec_in[i], dec_out[i], bilinear, kernels_per_layer=k))
        self.outc = OutConv(64, n_classes)

    def forward(self, x):
        """The reference's graph, block for block and in its call order (models/SmaAt_UNet.py:41-57): plain calls only --
        exactly what a ``patch_reference()`` user of the unchanged reference class executes.  The max-pool fusion still
        happens: ``cbamN(f)`` leaves MaxPool2d(2)(f) behind for the ``downN(f)`` that follows (modules.CBAM.forward)."""
        f = self.inc(x)
        att = [self.cbam1(f)]
        for lvl in range(1, 5):
            f = getattr(self, f"down{lvl}")(f)
            att.append(getattr(self, f"cbam{lvl + 1}")(f))
        y = att[4]                                                  # x5Att is the decoder input
        for i in range(4):
            y = getattr(self, f"up{i + 1}")(y, att[3 - i])          # attended maps are the skips
        return self.outc(y)

    def forward_serving(self, x):
        """Same graph with the one fusion the plain-call API cannot express: up4's last DS conv applies the 1-class OutConv
        in its epilogue (SmaAt_UNet.py:55-56), so the 64-channel activation never reaches HBM.  Used by
        ``engine.InferenceSession`` (inference only; falls back to the plain calls under autograd / train mode)."""
        f = self.inc(x)
        att = [self.cbam1(f)]
        for lvl in range(1, 5):
            f = getattr(self, f"down{lvl}")(f)
            att.append(getattr(self, f"cbam{lvl + 1}")(f))
        y = att[4]
        for i in range(3):
            y = getattr(self, f"up{i + 1}")(y, att[3 - i])
        return self.up4(y, att[0], outconv=self.outc)
