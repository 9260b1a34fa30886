"""CPU restatement of the reference's loss and metric bookkeeping -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.  numpy, float32 where the
reference computes in float32.  Follows:
  * UNetBase.loss_func                  reference models/regression_lightning.py:57-65
  * PrecipitationMetrics.update/compute reference metric/precipitation_metrics.py:37-95, :97-147
Pinned against the unmodified reference class (run under a minimal torchmetrics.Metric stand-in, torchmetrics is
not installed here) by oracle/make_golden_metrics.py -> tests/golden/precip_metrics.npz.
"""
from __future__ import annotations

import numpy as np

FACTOR = np.float32(47.83)  # precipitation_metrics.py:23


def loss_func(y_pred: np.ndarray, y_true: np.ndarray) -> np.float32:
    """regression_lightning.py:57-65: squeeze/unsqueeze dim 1 to match, then sum((p-y)^2) / B."""
    if y_pred.ndim > y_true.ndim:
        y_pred = np.squeeze(y_pred, 1)
    elif y_true.ndim > y_pred.ndim:
        y_pred = np.expand_dims(y_pred, 1)
    d = y_pred.astype(np.float32) - y_true.astype(np.float32)
    return np.float32(np.sum(d.astype(np.float64) ** 2) / y_true.shape[0])


def new_state() -> dict:
    """precipitation_metrics.py:26-34 (add_state defaults)."""
    return dict(total_loss=0.0, total_loss_denorm=0.0, total_samples=0, total_pixels=0,
                total_tp=0, total_fp=0, total_tn=0, total_fn=0)


def update(state: dict, preds: np.ndarray, target: np.ndarray, threshold: float = 0.5, denormalize: bool = True) -> dict:
    """precipitation_metrics.py:37-95."""
    preds = np.asarray(preds, np.float32)
    target = np.asarray(target, np.float32)
    if np.isnan(preds).any() or np.isnan(target).any():          # :46-48 batch ignored
        return state
    if preds.shape != target.shape:                               # :51-58
        if preds.ndim < target.ndim:
            preds = preds[None]
        elif preds.ndim > target.ndim:
            preds = np.squeeze(preds)
            if preds.ndim < target.ndim:
                preds = preds[None]
    bs = target.shape[0]                                          # :61
    d = preds - target
    state["total_loss"] += float(np.sum(d.astype(np.float64) ** 2) / bs)      # :62-63
    state["total_samples"] += bs                                  # :64
    state["total_pixels"] += target.size                          # :65
    if denormalize:                                               # :68-75
        pu, tu = preds * FACTOR, target * FACTOR
        dd = pu - tu
        state["total_loss_denorm"] += float(np.sum(dd.astype(np.float64) ** 2) / bs)
    else:
        pu, tu = preds, target
    pm = (pu * np.float32(12)) > np.float32(threshold)            # :80-85
    tm = (tu * np.float32(12)) > np.float32(threshold)
    conf = tm.reshape(-1).astype(np.int64) * 2 + pm.reshape(-1).astype(np.int64)   # :88
    bc = np.bincount(conf, minlength=4)                           # :89
    state["total_tn"] += int(bc[0]); state["total_fp"] += int(bc[1])               # :92-95
    state["total_fn"] += int(bc[2]); state["total_tp"] += int(bc[3])
    return state


def compute(state: dict, denormalize: bool = True) -> dict:
    """precipitation_metrics.py:97-147 (nan where the reference returns nan)."""
    nan = float("nan")
    tp, fp, tn, fn = (float(state[k]) for k in ("total_tp", "total_fp", "total_tn", "total_fn"))
    n = float(state["total_samples"])
    out = {}
    out["mse"] = state["total_loss"] / n if n else nan
    out["mse_denorm"] = state["total_loss_denorm"] / n if (denormalize and n) else nan
    out["mse_pixel"] = state["total_loss_denorm"] / state["total_pixels"] if (denormalize and state["total_pixels"]) else nan
    out["precision"] = tp / (tp + fp) if tp + fp > 0 else nan
    out["recall"] = tp / (tp + fn) if tp + fn > 0 else nan
    out["accuracy"] = (tp + tn) / (tp + tn + fp + fn) if (tp + tn + fp + fn) > 0 else nan
    p, r = out["precision"], out["recall"]
    out["f1"] = 2 * p * r / (p + r) if (p == p and r == r and p + r > 0) else nan
    out["csi"] = tp / (tp + fn + fp) if tp + fn + fp > 0 else nan
    out["far"] = fp / (tp + fp) if tp + fp > 0 else nan
    denom = (tp + fn) * (fn + tn) + (tp + fp) * (fp + tn)
    out["hss"] = (tp * tn - fn * fp) / denom if denom > 0 else nan
    return out


def metric_batches(seed: int = 0, n_batches: int = 4, B: int = 3, S: int = 24, nan_batch: int = 2):
    """Deterministic (preds[B,1,S,S], target[B,S,S]) batches shared by the golden generator and the tests.
    Values cluster around the 0.5 mm/h threshold (0.5/12/47.83 = 8.7e-4 normalised) so all four confusion cells
    fill; batch ``nan_batch`` carries one NaN and must be ignored."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_batches):
        t = (rng.random((B, S, S)) ** 4 * 4e-3).astype(np.float32)
        p = (t + rng.normal(0, 6e-4, (B, S, S))).astype(np.float32)[:, None]
        if i == nan_batch:
            p[1, 0, 3, 5] = np.nan
        out.append((p, t))
    return out
