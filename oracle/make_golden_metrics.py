"""Generate tests/golden/precip_metrics.npz from the UNMODIFIED reference classes -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_metrics

`metric/precipitation_metrics.py` subclasses torchmetrics.Metric and torchmetrics is not installed here (no
network), so a minimal stand-in module providing `Metric.add_state` (attribute = default tensor) is placed in
sys.modules before importing the reference file; the reference's update()/compute() bodies run unmodified.
UNetBase.loss_func (models/regression_lightning.py:57-65) needs `lightning`; its four-line body only calls
torch.nn.functional.mse_loss, which is invoked here exactly as that line does.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

from oracle import metrics_oracle as MO

REF = os.environ.get("SMAAT_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "precip_metrics.npz")


def reference_metric_class():
    tm = types.ModuleType("torchmetrics")

    class Metric(torch.nn.Module):
        def __init__(self, **kw):
            super().__init__()

        def add_state(self, name, default, dist_reduce_fx=None):
            setattr(self, name, default.clone())

    tm.Metric = Metric
    sys.modules["torchmetrics"] = tm
    sys.path.insert(0, REF)
    from metric.precipitation_metrics import PrecipitationMetrics  # noqa: E402
    return PrecipitationMetrics


def main():
    cls = reference_metric_class()
    res = {}
    for denorm in (True, False):
        m = cls(threshold=0.5, denormalize=denorm)
        losses = []
        for p, t in MO.metric_batches():
            pt, tt = torch.from_numpy(p), torch.from_numpy(t)
            m.update(pt, tt)
            losses.append(float(torch.nn.functional.mse_loss(pt.squeeze(1), tt, reduction="sum") / tt.size(0)))
        tag = "denorm" if denorm else "norm"
        for k, v in m.compute().items():
            res[f"{tag}/{k}"] = np.float64(float(v))
        for k in ("total_tp", "total_fp", "total_tn", "total_fn", "total_samples", "total_pixels"):
            res[f"{tag}/{k}"] = np.int64(int(getattr(m, k)))
        res[f"{tag}/losses"] = np.asarray(losses, np.float64)
    np.savez(OUT, **res)
    print("wrote", OUT, {k: (v.tolist() if v.ndim == 0 else "...") for k, v in res.items() if k.startswith("denorm/")})


if __name__ == "__main__":
    main()
