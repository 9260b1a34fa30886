"""Minimal stand-ins for the reference's missing third-party imports -- TEST INFRASTRUCTURE ONLY.

`models/unet_precip_regression_lightning.py` / `models/regression_lightning.py` import `lightning`, `torchmetrics`
(through metric/precipitation_metrics.py) and `h5py` (through utils/dataset_precip.py); none is installed here and
there is no network.  The stand-ins provide only what those files touch AT IMPORT / CONSTRUCTION time
(`pl.LightningModule` = nn.Module with `save_hyperparameters` / `log`, `torchmetrics.Metric.add_state`, an empty
`h5py`), so that the reference's own model classes -- constructor and `forward` bodies unmodified -- can be built and
run.  Nothing here computes anything.
"""
from __future__ import annotations

import argparse
import sys
import types

import torch


def install():
    if "lightning" not in sys.modules:
        lightning = types.ModuleType("lightning")
        pl = types.ModuleType("lightning.pytorch")

        class LightningModule(torch.nn.Module):
            def save_hyperparameters(self, hparams=None, *a, **kw):
                if isinstance(hparams, dict):
                    hparams = argparse.Namespace(**hparams)
                self.hparams = hparams

            def log(self, *a, **kw):
                pass

        pl.LightningModule = LightningModule
        lightning.pytorch = pl
        sys.modules["lightning"] = lightning
        sys.modules["lightning.pytorch"] = pl
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")

        class Metric(torch.nn.Module):
            def __init__(self, **kw):
                super().__init__()

            def add_state(self, name, default, dist_reduce_fx=None):
                setattr(self, name, default.clone())

        tm.Metric = Metric
        sys.modules["torchmetrics"] = tm
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")


def hparams(n_channels, n_classes, k, bilinear=True, reduction_ratio=16):
    """The hyper-parameters the Lightning wrappers read in their constructors (unet_precip_regression_lightning.py:122-128)."""
    return argparse.Namespace(n_channels=n_channels, n_classes=n_classes, kernels_per_layer=k, bilinear=bilinear,
                              reduction_ratio=reduction_ratio, learning_rate=1e-3, lr_patience=5, threshold=0.5)
