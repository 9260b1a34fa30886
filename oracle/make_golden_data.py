"""Generate tests/golden/dataset_semantics.npz from the UNMODIFIED reference dataset classes -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_data

`utils/dataset_precip.py` imports h5py (not installed here, no network); a minimal stand-in module whose
`File(name)[split]["images"]` returns seeded numpy arrays is placed in sys.modules, then the reference's
`precipitation_maps_oversampled_h5.__getitem__` / `precipitation_maps_h5.__getitem__` / `__len__` run unmodified
(utils/dataset_precip.py:6-77).  The arrays come from `data_arrays()` below, which the tests call as well.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF = os.environ.get("SMAAT_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dataset_semantics.npz")
INDICES = (0, 1, 7, 13)


def data_arrays():
    """(oversampled (N, T, H, W) with T = 12 inputs + 6 outputs, sequence (n_images, H, W)), deterministic float32."""
    rng = np.random.default_rng(20240924)
    over = rng.random((14, 18, 6, 5), dtype=np.float32)
    seq = rng.random((40, 6, 5), dtype=np.float32)
    return over, seq


def main():
    over, seq = data_arrays()
    files = {"over.h5": {"train": {"images": over}, "test": {"images": over[::-1].copy()}},
             "seq.h5": {"train": {"images": seq}, "test": {"images": seq[::-1].copy()}}}
    h5 = types.ModuleType("h5py")
    h5.File = lambda name, mode="r", **kw: files[name]
    sys.modules["h5py"] = h5
    sys.path.insert(0, REF)
    from utils import dataset_precip as D  # noqa: E402
    res = {}
    for train in (True, False):
        tag = "train" if train else "test"
        ds = D.precipitation_maps_oversampled_h5("over.h5", 12, 6, train=train)
        res[f"over/{tag}/len"] = np.int64(len(ds))
        for i in INDICES:
            x, y = ds[i]
            res[f"over/{tag}/{i}/x"], res[f"over/{tag}/{i}/y"] = x, y
        ds = D.precipitation_maps_h5("seq.h5", 12, 6, train=train)
        res[f"seq/{tag}/len"] = np.int64(len(ds))
        for i in INDICES:
            x, y = ds[i]
            res[f"seq/{tag}/{i}/x"], res[f"seq/{tag}/{i}/y"] = x, y
    np.savez(OUT, **res)
    print("wrote", OUT, len(res), "arrays")


if __name__ == "__main__":
    main()
