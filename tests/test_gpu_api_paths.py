"""-m gpu: the paths that are actually TIMED (engine.InferenceSession: CUDA-graph replay, pipelined submit/collect) and
ADVERTISED (the blocks called plainly, in the order the unchanged reference classes call them) against the oracle /
the golden outputs of the unmodified reference; one gradient check at BASELINE.json's real frame size; and the
train -> eval -> train -> eval cache-staleness sequence."""
import os

import numpy as np
import pytest
import torch
from torch import nn

import smaat_unet_b200 as S
from oracle import torch_port as TP
from oracle.cases import CASES, case_tensors, cast_sd, fill_schema, smaat_unet_schema
from smaat_unet_b200.engine import InferenceSession
from tests._util import NET_TOL, assert_close, dev, load_np_state_dict

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def make_model(seed=5):
    sd = cast_sd(fill_schema(smaat_unet_schema(12, 1, 2), seed), np.float32)
    m = load_np_state_dict(S.SmaAt_UNet(12, 1, kernels_per_layer=2), sd).cuda().eval()
    return m, TP.to_torch_sd(sd)


def _port(x, sd, frames):
    with torch.no_grad():
        return TP.smaat_unet_forward(x[frames], sd).double().numpy()


# ---------------------------------------------------------------------------------------------------------------------
# (a) what bench.py times: InferenceSession at B=32, 12x288x288
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("serving_fusions", [True, False])
def test_inference_session_b32_graph_and_pipeline_match_cpu_port(serving_fusions):
    m, sd = make_model()
    sess = InferenceSession(m, 32, (12, 288, 288), serving_fusions=serving_fusions)
    assert sess.graph is not None and sess.launches_per_forward > 40
    rng = np.random.default_rng(21)
    # device-resident path: graph replay on a fresh input (the captured warm-up ran on zeros)
    x = torch.from_numpy(rng.uniform(0, 1, (32, 12, 288, 288)).astype(np.float32))
    y = sess.forward(x.cuda()).clone()
    torch.cuda.synchronize()
    frames = [0, 13, 31]
    assert_close(y[frames], _port(x, sd, frames), NET_TOL["tf32x3"], "InferenceSession.forward B=32")
    y_again = sess.forward(x.cuda())
    torch.cuda.synchronize()
    assert torch.equal(y, y_again), "graph replay is not deterministic"
    # host-buffer path: 5 batches with DIFFERENT inputs through the 2-slot submit/collect pipeline (slots are reused
    # twice): a stale slot, a missing event or an overwritten staging buffer shows up as a wrong batch
    hosts = [torch.from_numpy(rng.uniform(0, 1, (32, 12, 288, 288)).astype(np.float32)).pin_memory() for _ in range(5)]
    outs = []
    sess.submit(hosts[0])
    for i in range(1, 5):
        sess.submit(hosts[i])
        outs.append(sess.collect().clone())
    outs.append(sess.collect().clone())
    for i, (h, o) in enumerate(zip(hosts, outs)):
        fr = [i, 31 - 3 * i]
        assert_close(o[fr], _port(h, sd, fr), NET_TOL["tf32x3"], f"submit/collect batch {i}")
    # and the whole batch must equal the device-resident path bit for bit (same graph, same data)
    y0 = sess.forward(hosts[0].cuda())
    torch.cuda.synchronize()
    assert torch.equal(y0.cpu(), outs[0])


def test_inference_session_refresh_follows_the_weights():
    """A session is a snapshot; refresh() must pick up weights written behind torch's back (raw-pointer kernels)."""
    m, sd = make_model(seed=6)
    sess = InferenceSession(m, 2, (12, 64, 64))
    x = torch.rand(2, 12, 64, 64, device="cuda")
    y0 = sess.forward(x).clone()
    with torch.no_grad():
        bn = m.inc.double_conv[1]
        bn.running_mean.add_(0.25)                  # any parameter change
    sess.refresh()
    y1 = sess.forward(x).clone()
    with torch.no_grad():
        y_eager = m(x)
    torch.cuda.synchronize()
    assert not torch.equal(y0, y1)
    assert_close(y1, y_eager.double().cpu().numpy(), 1e-6, "refreshed session vs eager")


# ---------------------------------------------------------------------------------------------------------------------
# (b) the advertised API: blocks called plainly, in the reference classes' own order
# ---------------------------------------------------------------------------------------------------------------------
class RefOrderNet(nn.Module):
    """Constructor and forward of the reference classes restated call for call with the drop-in blocks:
    models/SmaAt_UNet.py:23-57 and models/unet_precip_regression_lightning.py:130-164 (5 CBAMs),
    :175-208 (4 CBAMs, x5 goes to the decoder un-attended), :95-117 (UNetDS, none).  This is what a
    ``patch_reference()`` user of the unchanged reference code executes: plain ``cbam(x)``, ``down(x)``, ``up(a, b)``,
    ``outc(x)`` -- no keyword extensions."""

    def __init__(self, n_channels, n_classes, k, n_cbams):
        super().__init__()
        self.n_cbams = n_cbams
        self.inc = S.DoubleConvDS(n_channels, 64, kernels_per_layer=k)
        ch = [64, 128, 256, 512, 512]
        for i in range(1, 5):
            if i <= n_cbams:
                setattr(self, f"cbam{i}", S.CBAM(ch[i - 1], reduction_ratio=16))
            setattr(self, f"down{i}", S.DownDS(ch[i - 1], ch[i], kernels_per_layer=k))
        if n_cbams == 5:
            self.cbam5 = S.CBAM(512, reduction_ratio=16)
        self.up1 = S.UpDS(1024, 256, True, kernels_per_layer=k)
        self.up2 = S.UpDS(512, 128, True, kernels_per_layer=k)
        self.up3 = S.UpDS(256, 64, True, kernels_per_layer=k)
        self.up4 = S.UpDS(128, 64, True, kernels_per_layer=k)
        self.outc = S.OutConv(64, n_classes)

    def forward(self, x):
        att = (lambda i, t: getattr(self, f"cbam{i}")(t)) if self.n_cbams else (lambda i, t: t)
        x1 = self.inc(x)
        x1Att = att(1, x1)
        x2 = self.down1(x1)
        x2Att = att(2, x2)
        x3 = self.down2(x2)
        x3Att = att(3, x3)
        x4 = self.down3(x3)
        x4Att = att(4, x4)
        x5 = self.down4(x4)
        x5Att = self.cbam5(x5) if self.n_cbams == 5 else x5
        x = self.up1(x5Att, x4Att)
        x = self.up2(x, x3Att)
        x = self.up3(x, x2Att)
        x = self.up4(x, x1Att)
        return self.outc(x)


REF_ORDER_CASES = [("unet_12_1_k2_32", 5), ("unet_12_1_k2_odd", 5), ("unet_3_5_k1_48", 5), ("lit_dsatt_k2_32", 5),
                   ("lit_dsatt4_k2_48", 4), ("lit_ds_k1_32", 0)]


@pytest.mark.parametrize("mode", ["tf32x3", "tf32", "fp32"])
@pytest.mark.parametrize("name,n_cbams", REF_ORDER_CASES)
def test_reference_call_order_matches_reference_golden(name, n_cbams, mode):
    c = CASES[name]
    sd, xs = case_tensors(name, np.float32)
    net = load_np_state_dict(RefOrderNet(c["n_channels"], c["n_classes"], c["k"], n_cbams), sd).cuda().eval()
    S.set_pointwise_mode(mode)
    try:
        with torch.no_grad(), S.ops.profile() as prof:
            y = net(dev(xs[0]))
        names = [r[0].split("[")[0] for r in prof.records]
    finally:
        S.set_pointwise_mode("tf32x3")
    ref = np.load(os.path.join(GOLD, name + ".npz"))["output"]
    assert_close(y, ref, NET_TOL[mode], f"{name} reference call order [{mode}]")
    H, W = c["x"][2], c["x"][3]
    if n_cbams >= 4 and H % 16 == 0 and W % 32 == 0:
        # the plain calls reach the pool+max-pool fusion: no standalone max-pool launch is left
        assert names.count("smaat_maxpool2_fwd") == 0 and names.count("smaat_cbam_pool_mlp_fwd") >= 3, names
    if n_cbams == 0:
        assert names.count("smaat_maxpool2_fwd") == 4


def test_model_forward_is_the_reference_order_and_serving_forward_agrees():
    m, sd = make_model(seed=8)
    x = torch.from_numpy(np.random.default_rng(3).uniform(0, 1, (2, 12, 288, 288)).astype(np.float32))
    with torch.no_grad():
        ref = TP.smaat_unet_forward(x, sd).double().numpy()
        with S.ops.profile() as p1:
            y_plain = m(x.cuda())
        with S.ops.profile() as p2:
            y_serv = m.forward_serving(x.cuda())
    n1 = [r[0].split("[")[0] for r in p1.records]
    n2 = [r[0].split("[")[0] for r in p2.records]
    assert_close(y_plain, ref, NET_TOL["tf32x3"], "plain forward")
    assert_close(y_serv, ref, NET_TOL["tf32x3"], "serving forward")
    assert "smaat_maxpool2_fwd" not in n1 and n1.count("smaat_cbam_pool_mlp_fwd") == 3 and n1.count("smaat_cbam_gate_scale_fwd") == 4   # cbam5 is 18 x 18: W % 4 != 0 -> gate + scale kernels
    assert n1.count("smaat_cbam_mlp_fwd") == 2 and n1.count("smaat_cbam_pool_maxpool_fwd") == 1    # the two 512-channel CBAMs: pool and MLP as two launches (faster)
    assert "smaat_outconv_fwd" in n1 and "smaat_outconv_fwd" not in n2 and "smaat_dsconv_outconv_fwd" in n2


def test_maxpool_stash_is_only_taken_for_the_same_tensor():
    torch.manual_seed(0)
    cb = S.CBAM(32).cuda().eval()
    dn = S.DownDS(32, 64, kernels_per_layer=2).cuda().eval()
    x = torch.randn(2, 32, 16, 16, device="cuda")
    other = torch.randn(2, 32, 16, 16, device="cuda")
    with torch.no_grad():
        want = dn(x.clone())                        # no stash: own max-pool kernel
        want_other = dn(other.clone())
        cb(x)
        with S.ops.profile() as p:
            got_other = dn(other)                   # different tensor: must NOT take x's max-pool
        assert any(r[0].startswith("smaat_maxpool2_fwd") for r in p.records)
        with S.ops.profile() as p:
            got = dn(x)                             # same tensor: takes it
        assert not any(r[0].startswith("smaat_maxpool2_fwd") for r in p.records)
        cb(x)
        x.mul_(2.0)                                 # in-place change after the CBAM: version differs -> recompute
        got2 = dn(x)
        want2 = dn(x.clone())
    assert torch.equal(got_other, want_other) and torch.equal(got, want) and torch.equal(got2, want2)


# ---------------------------------------------------------------------------------------------------------------------
# (c) gradients at BASELINE.json's frame size (configs[2] geometry, B=2): split-K wgrad atomics over P = 82 944,
#     fp64 BatchNorm sums over n = 165 888, vs float64 CPU autograd over the port
# ---------------------------------------------------------------------------------------------------------------------
def _port_train_step(sd_np, x_np, t_np, dtype):
    """fwd + loss_func (regression_lightning.py:57-65) + backward over the CPU port in ``dtype``; parameters hold the fp32-rounded
    values the GPU model holds.  Returns (loss, y, {name: grad}, dL/dx) as float64 numpy."""
    sd = {}
    for k, v in sd_np.items():
        t = torch.as_tensor(np.asarray(v))
        if t.dtype == torch.int64:
            sd[k] = t
        elif k.endswith(("running_mean", "running_var")):
            sd[k] = t.float().to(dtype)
        else:
            sd[k] = t.float().to(dtype).requires_grad_(True)
    xr = torch.from_numpy(x_np).float().to(dtype).requires_grad_(True)
    yr = TP.smaat_unet_forward(xr, sd, True)
    loss = torch.nn.functional.mse_loss(yr.squeeze(1), torch.from_numpy(t_np).float().to(dtype), reduction="sum") / x_np.shape[0]
    loss.backward()
    grads = {k: v.grad.double().numpy() for k, v in sd.items() if v.dtype != torch.int64 and v.requires_grad}
    return float(loss.detach()), yr.detach().double().numpy(), grads, xr.grad.double().numpy()


def test_gradients_at_288x288_match_cpu_autograd_fp64():
    seed = 12
    sd_np = fill_schema(smaat_unet_schema(12, 1, 2), seed)
    m = load_np_state_dict(S.SmaAt_UNet(12, 1, kernels_per_layer=2), cast_sd(sd_np, np.float32)).cuda().train()
    rng = np.random.default_rng(4)
    x_np = rng.uniform(0, 1, (2, 12, 288, 288))
    t_np = rng.uniform(0, 1, (2, 288, 288))
    x = dev(x_np).requires_grad_(True)
    y = m(x)
    loss = torch.nn.functional.mse_loss(y.squeeze(1), dev(t_np), reduction="sum") / 2
    loss.backward()
    torch.cuda.synchronize()
    l64, y64, g64, x64 = _port_train_step(sd_np, x_np, t_np, torch.float64)      # ground truth
    l32, y32, g32, x32 = _port_train_step(sd_np, x_np, t_np, torch.float32)      # the reference algorithm's own fp32 movement
    assert abs(float(loss) - l64) <= 2e-4 * abs(l64)
    assert_close(y, y64, 3e-4, "288x288 train forward")

    def rel_max(a, b):
        return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

    def rel_l2(a, b):
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

    gmax = max(float(np.abs(g).max()) for g in g64.values())
    live = [k for k, g in g64.items() if np.abs(g).max() >= 1e-6 * gmax]
    # With B=2 the 18x18 levels normalise over n = 648 values per channel and the reference itself moves by ~1e-2 (max norm) /
    # ~3e-3 (L2) between fp32 and fp64 here (measured: CPU port, this seed).  One noise level for the whole case; the CUDA
    # path must stay within a small multiple of it.
    noise_max = max(rel_max(g32[k], g64[k]) for k in live)
    noise_l2 = max(rel_l2(g32[k], g64[k]) for k in live)
    tol_max, tol_l2 = max(2e-3, 5.0 * noise_max), max(1e-3, 5.0 * noise_l2)
    worst = ("", 0.0)
    for name, p in m.named_parameters():
        got = p.grad.double().cpu().numpy()
        if name not in live:      # mathematically zero (conv bias cancelled by the batch-mean subtraction): only summation noise
            assert float(np.abs(got).max()) <= 1e-3 * gmax, name
            continue
        e_max, e_l2 = rel_max(got, g64[name]), rel_l2(got, g64[name])
        if e_max > worst[1]:
            worst = (name, e_max)
        assert np.isfinite(e_max) and e_max <= tol_max and e_l2 <= tol_l2, \
            f"grad {name}: rel max {e_max:.3e} (tol {tol_max:.1e}), rel L2 {e_l2:.3e} (tol {tol_l2:.1e})"
    # dL/dx is per pixel: a max-pool / channel-max argmax that flips under fp32 noise reroutes one pixel's gradient (the port in
    # fp32 is off by 3e-2 of the maximum at isolated pixels), so the input gradient is held to noise-calibrated relative
    # L2 and max-norm bounds
    gx = x.grad.double().cpu().numpy()
    e_l2, e_max, n_l2, n_max = rel_l2(gx, x64), rel_max(gx, x64), rel_l2(x32, x64), rel_max(x32, x64)
    print(f"dL/dx: rel L2 {e_l2:.2e} (reference fp32: {n_l2:.2e}), rel max {e_max:.2e} (reference fp32: {n_max:.2e})")
    assert e_l2 <= max(2e-3, 3.0 * n_l2) and e_max <= max(2e-3, 5.0 * n_max), \
        f"dL/dx: rel L2 {e_l2:.3e} vs port fp32 {n_l2:.3e}; rel max {e_max:.3e} vs {n_max:.3e}"
    print(f"worst parameter gradient: {worst[0]} rel max {worst[1]:.2e}; reference fp32 noise: max {noise_max:.2e}, L2 {noise_l2:.2e}")


# ---------------------------------------------------------------------------------------------------------------------
# (d) ADVICE r1 (high): caches derived from weights must not survive raw-pointer / graph-replay updates
# ---------------------------------------------------------------------------------------------------------------------
def test_train_eval_train_eval_uses_current_weights():
    from smaat_unet_b200.train import TrainSession
    torch.manual_seed(0)
    m = S.SmaAt_UNet(12, 1, kernels_per_layer=2).cuda()
    sess = TrainSession(m, batch=2, in_shape=(12, 64, 64), lr=1e-2, use_graph=True)
    g = torch.Generator(device="cuda").manual_seed(1)
    xv = torch.rand(2, 12, 64, 64, device="cuda", generator=g)

    def eval_now():
        m.eval()
        with torch.no_grad():
            y = m(xv).clone()
        m.train()
        return y

    def eval_fresh():
        """Same weights in a freshly built model: no cache can be stale."""
        m2 = S.SmaAt_UNet(12, 1, kernels_per_layer=2).cuda()
        m2.load_state_dict(m.state_dict())
        m2.eval()
        with torch.no_grad():
            return m2(xv).clone()

    for rnd in range(2):
        for _ in range(3):
            sess.step(torch.rand(2, 12, 64, 64, device="cuda", generator=g), torch.rand(2, 64, 64, device="cuda", generator=g))
        a, b = eval_now(), eval_fresh()
        torch.cuda.synchronize()
        assert torch.equal(a, b), f"round {rnd}: eval after graph-replayed training used stale folded weights"
    # eager train-mode forward WITHOUT an optimizer step: running statistics move by raw pointer only
    m.eval()
    with torch.no_grad():
        m(xv)                                        # fills the caches
    m.train()
    with torch.no_grad():
        m(xv * 2.0)                                  # updates running stats (no _version bump)
    assert torch.equal(eval_now(), eval_fresh())
