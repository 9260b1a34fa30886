"""Loss + PrecipitationMetrics step: oracle vs the reference's own outputs (CPU), CUDA vs oracle (-m gpu)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "precip_metrics.npz"))
KEYS = ("mse", "mse_denorm", "mse_pixel", "precision", "recall", "accuracy", "f1", "csi", "far", "hss")
COUNTS = ("total_tp", "total_fp", "total_tn", "total_fn", "total_samples", "total_pixels")


def _same(a, b, tol, what):
    a, b = float(a), float(b)
    if math.isnan(b):
        assert math.isnan(a), what
    else:
        assert abs(a - b) <= tol * max(abs(b), 1e-30), f"{what}: {a} vs {b}"


@pytest.mark.parametrize("denorm", [True, False])
def test_oracle_matches_reference_metrics(denorm):
    tag = "denorm" if denorm else "norm"
    st = MO.new_state()
    losses = []
    for p, t in MO.metric_batches():
        MO.update(st, p, t, 0.5, denorm)
        losses.append(MO.loss_func(p, t))
    for k in COUNTS:                                    # integer states: bit-exact
        assert int(st[k]) == int(GOLD[f"{tag}/{k}"]), k
    out = MO.compute(st, denorm)
    for k in KEYS:                                      # the reference accumulates/divides in float32
        _same(out[k], GOLD[f"{tag}/{k}"], 2e-6, f"{tag}/{k}")
    ref_l = GOLD[f"{tag}/losses"]
    for i, l in enumerate(losses):
        _same(l, ref_l[i], 2e-6, f"loss[{i}]")


def test_nan_batch_is_ignored_by_the_oracle():
    st = MO.new_state()
    p, t = MO.metric_batches()[2]
    assert np.isnan(p).any()
    MO.update(st, p, t)
    assert st == MO.new_state()


# ----------------------------------------------------------------------------------------------- gpu
@pytest.mark.gpu
@pytest.mark.parametrize("denorm", [True, False])
def test_cuda_metrics_match_oracle_and_reference(denorm):
    import smaat_unet_b200 as S
    tag = "denorm" if denorm else "norm"
    m = S.PrecipitationMetrics(threshold=0.5, denormalize=denorm)
    st = MO.new_state()
    for p, t in MO.metric_batches():
        m.update(torch.from_numpy(p).cuda(), torch.from_numpy(t).cuda())
        MO.update(st, p, t, 0.5, denorm)
    for k in COUNTS:
        assert int(getattr(m, k)) == int(st[k]) == int(GOLD[f"{tag}/{k}"]), k
    assert int(m.skipped_batches) == 1
    got, want = m.compute(), MO.compute(st, denorm)
    for k in KEYS:
        _same(got[k], want[k], 1e-6, f"{tag}/{k} vs oracle")
        _same(got[k], GOLD[f"{tag}/{k}"], 2e-6, f"{tag}/{k} vs reference")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 24, 24), (2, 17, 19), (32, 288, 288)])
def test_cuda_loss_and_gradient(shape):
    import smaat_unet_b200 as S
    g = torch.Generator().manual_seed(sum(shape))
    t = torch.rand(shape, generator=g)
    p = (t + 0.1 * torch.randn(shape, generator=g)).unsqueeze(1)
    pc = p.cuda().requires_grad_(True)
    m = S.PrecipitationMetrics()
    loss = S.step_loss(pc, t.cuda(), m)
    (loss * 3.0).backward()
    p64 = p.double().requires_grad_(True)
    ref = torch.nn.functional.mse_loss(p64.squeeze(1), t.double(), reduction="sum") / shape[0]
    (ref * 3.0).backward()
    _same(loss.item(), ref.item(), 1e-6, "loss")
    assert float(loss.item()) == pytest.approx(float(MO.loss_func(p.numpy(), t.numpy())), rel=1e-6)
    err = (pc.grad.cpu().double() - p64.grad).abs().max().item() / p64.grad.abs().max().item()
    assert err <= 1e-6, err
    st = MO.update(MO.new_state(), p.numpy(), t.numpy())
    for k in COUNTS:                                    # full-size confusion counts: bit-exact vs the oracle
        assert int(getattr(m, k)) == int(st[k]), k
    # size-independent property: TN+FP+FN+TP == pixels
    assert int(m.total_tn + m.total_fp + m.total_fn + m.total_tp) == t.numel()


@pytest.mark.gpu
def test_cuda_loss_func_unaligned_and_no_grad():
    import smaat_unet_b200 as S
    base_p = torch.rand(1 + 2 * 5 * 7, device="cuda")
    base_t = torch.rand(1 + 2 * 5 * 7, device="cuda")
    p, t = base_p[1:].view(2, 1, 5, 7), base_t[1:].view(2, 5, 7)      # 4-byte aligned only -> scalar kernel
    loss = S.loss_func(p, t)
    ref = torch.nn.functional.mse_loss(p.squeeze(1).double(), t.double(), reduction="sum") / 2
    _same(loss.item(), ref.item(), 1e-6, "loss")
