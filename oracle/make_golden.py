"""Generate tests/golden/*.npz by running the UNMODIFIED reference -- TEST INFRASTRUCTURE ONLY.

Run in the build container (the only place /root/reference exists):

    python -m oracle.make_golden

For every case in ``oracle/cases.py`` this builds the reference module from
/root/reference (models/layers.py, models/unet_parts_depthwise_separable.py,
models/unet_parts.py, models/SmaAt_UNet.py), checks that its ``state_dict()``
keys and shapes equal the schema restated in cases.py, loads the deterministic
float64 parameters, runs the reference forward in float64 on the deterministic
input, and stores the output (and, for train-mode cases, the BatchNorm buffers
after the step).  The reference is torch==2.6.0-pinned; this container runs
torch 2.11 (version skew recorded in the fixture's ``meta``).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SMAAT_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_reference_module(c):
    sys.path.insert(0, REF)
    from models.layers import CBAM, DepthwiseSeparableConv          # noqa: E402
    from models.SmaAt_UNet import SmaAt_UNet                         # noqa: E402
    from models.unet_parts import OutConv                            # noqa: E402
    from models.unet_parts_depthwise_separable import DoubleConvDS, DownDS, UpDS  # noqa: E402

    kind = c["kind"]

    class Wrap(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, *a):
            return self.m(*a)

    if kind == "dsconv":
        return Wrap(DepthwiseSeparableConv(c["cin"], c["cout"], kernel_size=3, padding=1, kernels_per_layer=c["k"]))
    if kind == "doubleconv":
        return Wrap(DoubleConvDS(c["cin"], c["cout"], c["mid"], kernels_per_layer=c["k"]))
    if kind == "down":
        return Wrap(DownDS(c["cin"], c["cout"], kernels_per_layer=c["k"]))
    if kind == "up":
        return Wrap(UpDS(c["cin"], c["cout"], c.get("bilinear", True), kernels_per_layer=c["k"]))
    if kind == "cbam":
        return Wrap(CBAM(c["c"], reduction_ratio=c["r"], kernel_size=c["ks"]))
    if kind == "outconv":
        return Wrap(OutConv(c["cin"], c["cout"]))
    if kind == "config1":
        class Block(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.conv = DoubleConvDS(c["c"], c["c"], kernels_per_layer=c["k"])
                self.cbam = CBAM(c["c"])

            def forward(self, x):
                return self.cbam(self.conv(x))
        return Block()
    if kind == "unet":
        return SmaAt_UNet(c["n_channels"], c["n_classes"], kernels_per_layer=c["k"])
    if kind == "lit":
        # the Lightning wrapper classes themselves, constructor and forward unmodified, under stand-ins for the
        # imports that are not installed here (oracle/ref_stubs.py)
        from oracle import ref_stubs
        ref_stubs.install()
        import models.unet_precip_regression_lightning as L          # noqa: E402
        return getattr(L, c["cls"])(hparams=ref_stubs.hparams(c["n_channels"], c["n_classes"], c["k"]))
    raise KeyError(kind)


def main():
    from oracle.cases import CASES, case_schema, case_tensors
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    index = {}
    only = set(sys.argv[1:])          # optional: regenerate only the named cases (index.json is always rewritten in full)
    for name, c in CASES.items():
        mod = build_reference_module(c).double()
        ref_sd = mod.state_dict()
        schema = case_schema(c)
        assert set(ref_sd) == set(schema), (name, set(ref_sd) ^ set(schema))
        for k, v in ref_sd.items():
            assert tuple(v.shape) == tuple(schema[k]), (name, k, tuple(v.shape), schema[k])
        sd, xs = case_tensors(name, np.float64)
        mod.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        train = c.get("train", False)
        mod.train(train)
        with torch.no_grad():
            y = mod(*[torch.from_numpy(x) for x in xs])
        index[name] = {"output_shape": list(y.shape), "train": train,
                       "abs_max": float(y.abs().max()), "n_params": int(sum(int(np.prod(s)) for s in schema.values()))}
        if only and name not in only:
            continue
        store = np.float32 if c.get("store") == "f4" else np.float64
        arrays = {"output": y.numpy().astype(store)}
        if train:
            after = mod.state_dict()
            for k, v in after.items():
                if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
                    arrays["buf:" + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        print(f"{name:28s} out={tuple(y.shape)} absmax={index[name]['abs_max']:.4f}")
    meta = {"reference": "HansBambel/SmaAt-UNet @ /root/reference", "torch": torch.__version__,
            "reference_pins_torch": "2.6.0", "dtype": "float64 (config1_block stored as float32)", "cases": index}
    with open(os.path.join(OUT, "index.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
