"""smaat_unet_b200 -- B200 (sm_100a) implementation of the SmaAt-UNet forward hot path.

Drop-in ``nn.Module`` replacements for the reference's DS-conv blocks and CBAM
(``modules``), the same model assembly (``model.SmaAt_UNet``), a helper that rebinds the
reference's own classes (``patch_reference``), and the functional kernel wrappers (``ops``).
All arithmetic runs in ``libsmaat_b200.so`` (C ABI: ``include/smaat_b200.h``).
"""
from . import _lib, ops  # noqa: F401
from .data import PinnedBatchLoader, precipitation_maps_oversampled_shard, precipitation_maps_shard  # noqa: F401
from .metrics import PrecipitationMetrics, loss_func, step_loss  # noqa: F401
from .model import SmaAt_UNet  # noqa: F401
from .modules import (CBAM, ChannelAttention, DepthwiseSeparableConv, DoubleConvDS, DownDS, OutConv,  # noqa: F401
                      SpatialAttention, UpDS)
from .ops import get_pointwise_mode, set_fused_dsconv, set_pointwise_mode  # noqa: F401
from .patch import patch_reference  # noqa: F401

__all__ = ["SmaAt_UNet", "CBAM", "ChannelAttention", "SpatialAttention", "DepthwiseSeparableConv", "DoubleConvDS",
           "DownDS", "UpDS", "OutConv", "patch_reference", "PrecipitationMetrics", "loss_func", "step_loss", "set_pointwise_mode", "get_pointwise_mode", "ops"]
