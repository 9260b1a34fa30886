"""The oracle is only trustworthy once pinned: numpy oracle and torch port vs the
golden outputs produced by the UNMODIFIED reference (oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import torch_port as TP
from oracle.cases import CASES, case_schema, case_tensors, run_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def test_index_lists_every_case():
    with open(os.path.join(GOLD, "index.json")) as f:
        idx = json.load(f)
    assert set(idx["cases"]) == set(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_numpy_oracle_fp64_matches_reference(name):
    y, upd = run_oracle(name, np.float64)
    g = _gold(name)
    ref = g["output"].astype(np.float64)
    tol = 2e-7 if CASES[name].get("store") == "f4" else 1e-11   # f4 fixture: rounding of the stored value
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= tol * max(1.0, np.abs(ref).max())
    for k in g.files:
        if k.startswith("buf:"):
            got = np.asarray(upd[k[4:]], dtype=np.float64)
            assert np.abs(got - g[k]).max() <= 1e-12, k


@pytest.mark.parametrize("name", sorted(CASES))
def test_numpy_oracle_fp32_within_fp32_noise(name):
    """fp32 evaluation of the same algorithm: sets the noise floor the CUDA path is held to."""
    y, _ = run_oracle(name, np.float32)
    ref = _gold(name)["output"].astype(np.float64)
    tol = 2e-4 if CASES[name].get("train") else 2e-5
    assert np.abs(y - ref).max() <= tol * max(1.0, np.abs(ref).max())


_RUN = {
    "dsconv": lambda c, sd, xs: TP.ds_conv(xs[0], sd, "m"),
    "doubleconv": lambda c, sd, xs: TP.double_conv_ds(xs[0], sd, "m", c.get("train", False)),
    "down": lambda c, sd, xs: TP.down_ds(xs[0], sd, "m", c.get("train", False)),
    "up": lambda c, sd, xs: TP.up_ds(xs[0], xs[1], sd, "m", c.get("train", False)),
    "cbam": lambda c, sd, xs: TP.cbam(xs[0], sd, "m", c.get("train", False)),
    "outconv": lambda c, sd, xs: torch.nn.functional.conv2d(xs[0], sd["m.conv.weight"], sd["m.conv.bias"]),
    "config1": lambda c, sd, xs: TP.cbam(TP.double_conv_ds(xs[0], sd, "conv"), sd, "cbam"),
    "unet": lambda c, sd, xs: TP.smaat_unet_forward(xs[0], sd, c.get("train", False)),
    "lit": lambda c, sd, xs: TP.smaat_unet_forward(xs[0], sd, c.get("train", False), n_cbams=c["n_cbams"]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_torch_port_fp64_matches_reference(name):
    c = CASES[name]
    sd_np, xs_np = case_tensors(name, np.float64)
    sd = TP.to_torch_sd(sd_np, torch.float64)
    xs = [torch.from_numpy(x) for x in xs_np]
    with torch.no_grad():
        y = _RUN[c["kind"]](c, sd, xs).numpy()
    g = _gold(name)
    ref = g["output"].astype(np.float64)
    tol = 2e-7 if c.get("store") == "f4" else 1e-11
    assert np.abs(y - ref).max() <= tol * max(1.0, np.abs(ref).max())
    for k in g.files:                       # F.batch_norm updates running stats in place
        if k.startswith("buf:") and not k.endswith("num_batches_tracked"):
            assert np.abs(sd[k[4:]].numpy() - g[k]).max() <= 1e-12, k


def test_schema_param_count_matches_survey():
    # SmaAt_UNet(12,1,kpl=2) has 4 033 537 trainable parameters, 214 state_dict entries
    s = case_schema(CASES["unet_12_1_k2_32"])
    assert len(s) == 214
    n = sum(int(np.prod(v)) for k, v in s.items()
            if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert n == 4_033_537
