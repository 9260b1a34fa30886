"""-m gpu: the drop-in modules (called through the C ABI) vs the committed golden outputs of the
unmodified reference, for every eval-mode case, in every pointwise arithmetic mode."""
import os

import numpy as np
import pytest
import torch

import smaat_unet_b200 as S
from oracle.cases import CASES, case_tensors
from tests._util import NET_TOL, PW_TOL, assert_close, dev, load_np_state_dict

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
# the Lightning-wrapper cases ("lit") are run in the wrappers' own call order by tests/test_gpu_api_paths.py
EVAL_CASES = sorted(n for n, c in CASES.items() if not c.get("train", False) and c["kind"] != "lit")


def build(c):
    kind = c["kind"]
    if kind == "dsconv":
        return S.DepthwiseSeparableConv(c["cin"], c["cout"], 3, padding=1, kernels_per_layer=c["k"]), "m."
    if kind == "doubleconv":
        return S.DoubleConvDS(c["cin"], c["cout"], c["mid"], kernels_per_layer=c["k"]), "m."
    if kind == "down":
        return S.DownDS(c["cin"], c["cout"], kernels_per_layer=c["k"]), "m."
    if kind == "up":
        return S.UpDS(c["cin"], c["cout"], c.get("bilinear", True), kernels_per_layer=c["k"]), "m."
    if kind == "cbam":
        return S.CBAM(c["c"], reduction_ratio=c["r"], kernel_size=c["ks"]), "m."
    if kind == "outconv":
        return S.OutConv(c["cin"], c["cout"]), "m."
    if kind == "config1":
        class Block(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.conv = S.DoubleConvDS(c["c"], c["c"], kernels_per_layer=c["k"])
                self.cbam = S.CBAM(c["c"])

            def forward(self, x):
                return self.cbam(self.conv(x))
        return Block(), ""
    if kind == "unet":
        return S.SmaAt_UNet(c["n_channels"], c["n_classes"], kernels_per_layer=c["k"]), ""
    raise KeyError(kind)


@pytest.mark.parametrize("mode", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("name", EVAL_CASES)
def test_module_matches_reference_golden(name, mode):
    c = CASES[name]
    sd, xs = case_tensors(name, np.float32)
    mod, prefix = build(c)
    load_np_state_dict(mod, sd, prefix)
    mod = mod.cuda().eval()
    S.set_pointwise_mode(mode)
    try:
        with torch.no_grad():
            y = mod(*[dev(x) for x in xs])
        torch.cuda.synchronize()
    finally:
        S.set_pointwise_mode("tf32x3")
    ref = np.load(os.path.join(GOLD, name + ".npz"))["output"]
    tol = NET_TOL[mode] if c["kind"] == "unet" else PW_TOL[mode] * 2
    assert_close(y, ref, tol, f"{name} [{mode}]")


def test_standalone_attention_modules():
    from oracle import smaat_oracle as O
    name = "cbam_k7_eval"
    sd, xs = case_tensors(name, np.float32)
    m = load_np_state_dict(S.CBAM(32, 16, 7), sd, "m.").cuda().eval()
    x = dev(xs[0])
    sd64, xs64 = case_tensors(name, np.float64)
    with torch.no_grad():
        assert_close(m.channel_att(x), O.channel_attention(xs64[0], sd64, "m.channel_att"), 1e-5, "ChannelAttention")
        assert_close(m.spatial_att(x), O.spatial_attention(xs64[0], sd64, "m.spatial_att")[0], 1e-5, "SpatialAttention")


def test_unsupported_requests_are_refused_loudly_not_silently_wrong():
    up = S.UpDS(8, 4, bilinear=False).cuda().eval()   # only the reference's ConvTranspose2d(kernel_size=2, stride=2) exists (parts_ds.py:72)
    up.up = torch.nn.ConvTranspose2d(8, 4, kernel_size=3, stride=2).cuda()
    with pytest.raises(NotImplementedError), torch.no_grad():
        up(torch.zeros(1, 8, 4, 4, device="cuda"), torch.zeros(1, 4, 8, 8, device="cuda"))
    with pytest.raises(NotImplementedError):          # standalone attention halves have no autograd node (CBAM does)
        S.ChannelAttention(16).cuda()(torch.zeros(1, 16, 8, 8, device="cuda"))
    with pytest.raises(NotImplementedError):          # only the reference's 3x3 / padding=1 depthwise exists
        with torch.no_grad():
            S.DepthwiseSeparableConv(4, 8, 5, padding=2).cuda()(torch.zeros(1, 4, 8, 8, device="cuda"))
