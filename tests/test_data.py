"""Input pipeline: shard datasets vs the reference dataset classes' own outputs; loader ordering / sharding."""
import os

import numpy as np
import pytest
import torch

from oracle.make_golden_data import INDICES, data_arrays
from smaat_unet_b200 import data as D

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "dataset_semantics.npz"))


@pytest.mark.parametrize("train", [True, False])
def test_shard_datasets_reproduce_the_reference_samples(tmp_path, train):
    over, seq = data_arrays()
    tag = "train" if train else "test"
    over_s = over if train else over[::-1].copy()
    seq_s = seq if train else seq[::-1].copy()
    po = D.write_shard(str(tmp_path / "over.npy"), over_s)        # through the file (memory-mapped) ...
    ds_o = D.precipitation_maps_oversampled_shard(po, 12, 6, train=train)
    ds_s = D.precipitation_maps_shard(seq_s, 12, 6, train=train)  # ... and straight from an array
    assert len(ds_o) == int(GOLD[f"over/{tag}/len"]) and len(ds_s) == int(GOLD[f"seq/{tag}/len"])
    for i in INDICES:
        for name, ds in (("over", ds_o), ("seq", ds_s)):
            x, y = ds[i]
            assert x.dtype == np.float32 and np.array_equal(x, GOLD[f"{name}/{tag}/{i}/x"]) and np.array_equal(y, GOLD[f"{name}/{tag}/{i}/y"])
            xs, ys = ds.sample_shapes()
            xb, yb = np.empty(xs, np.float32), np.empty(ys, np.float32)
            ds.read_into(i, xb, yb)
            assert np.array_equal(xb, x) and np.array_equal(yb, y)


def test_transform_is_applied_like_the_reference():
    over, _ = data_arrays()
    ds = D.precipitation_maps_oversampled_shard(over, 12, 6, transform=lambda a: a * 2.0)
    x, y = ds[3]
    assert np.array_equal(x, over[3][:12] * 2.0) and np.array_equal(y, over[3][-1] * 2.0)
    xb, yb = np.empty((12, 6, 5), np.float32), np.empty((6, 5), np.float32)
    ds.read_into(3, xb, yb)
    assert np.array_equal(xb, x) and np.array_equal(yb, y)


class _Ev:
    def __init__(self):
        self.n = 0

    def synchronize(self):
        self.n += 1


def test_loader_order_batches_and_guard():
    over, _ = data_arrays()
    ds = D.precipitation_maps_oversampled_shard(over, 12, 6)
    ld = D.PinnedBatchLoader(ds, batch_size=4, drop_last=False, ring=2, pin_memory=False)
    assert len(ld) == 4
    ev, seen = _Ev(), []
    for x, y in ld:
        seen.append((x.clone(), y.clone()))
        ld.guard(ev)
    assert [len(x) for x, _ in seen] == [4, 4, 4, 2]
    assert torch.equal(torch.cat([x for x, _ in seen]), torch.from_numpy(over[:, :12])) and \
        torch.equal(torch.cat([y for _, y in seen]), torch.from_numpy(over[:, -1]))
    assert ev.n >= 2          # slots were re-used only after their guard event was waited for
    assert len(D.PinnedBatchLoader(ds, batch_size=4, drop_last=True, pin_memory=False)) == 3


def test_loader_shuffle_is_seeded_and_epoch_dependent():
    over, _ = data_arrays()
    ds = D.precipitation_maps_oversampled_shard(over, 12, 6)
    mk = lambda: D.PinnedBatchLoader(ds, batch_size=7, shuffle=True, seed=5, pin_memory=False)
    a, b = mk(), mk()
    assert a.epoch_indices() == b.epoch_indices() and sorted(a.epoch_indices()) == list(range(14))
    b.set_epoch(1)
    assert a.epoch_indices() != b.epoch_indices()
    ya = torch.cat([y.clone() for _, y in a])
    assert torch.equal(ya, torch.from_numpy(over[a.epoch_indices(), -1]))


@pytest.mark.parametrize("n,world", [(14, 2), (13, 4), (5, 8)])
def test_rank_sharding_matches_distributed_sampler(n, world):
    from torch.utils.data.distributed import DistributedSampler
    idx = list(range(n))
    for drop in (False, True):
        for r in range(world):
            want = list(DistributedSampler(idx, num_replicas=world, rank=r, shuffle=False, drop_last=drop))
            assert D.shard_indices(idx, r, world, drop_last=drop) == want, (r, drop)


def test_loader_explicit_index_list_like_the_reference_split():
    # regression_lightning.py:166-175: a shuffled index list split into train / valid subsets
    over, _ = data_arrays()
    ds = D.precipitation_maps_oversampled_shard(over, 12, 6)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(ds)).tolist()
    valid, train = perm[:3], perm[3:]
    got = torch.cat([y.clone() for _, y in D.PinnedBatchLoader(ds, batch_size=3, indices=valid, pin_memory=False)])
    assert torch.equal(got, torch.from_numpy(over[valid, -1]))
    halves = [D.PinnedBatchLoader(ds, batch_size=2, indices=train, rank=r, world=2, drop_last=False, pin_memory=False) for r in range(2)]
    assert sorted(set(halves[0].epoch_indices()) | set(halves[1].epoch_indices())) == sorted(train)
