"""world_size-2 gloo tests (CPU) of the multi-process host logic used by bench.py --gpus N."""
import os

import pytest
import torch
import torch.multiprocessing as mp

from smaat_unet_b200 import parallel as P


def test_shard_range_covers_batch_exactly():
    for n in (1, 7, 32, 33, 256):
        for world in (1, 2, 3, 8):
            spans = [P.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = P.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    P.barrier()
    slow = P.reduce_max(1.0 + rank)                     # timing reduction: the slowest rank wins
    lo, hi = P.shard_range(32, rank, world)             # batch sharding: disjoint, complete
    owned = torch.zeros(32)
    owned[lo:hi] = 1
    torch.distributed.all_reduce(owned)
    g1, g2 = torch.full((5,), float(rank + 1)), torch.full((2, 3), 10.0 * (rank + 1))
    n = P.allreduce_flat_([g1, None, g2], average=True)  # gradient bucket: mean over ranks
    q.put((rank, slow, owned.tolist(), n, g1.tolist(), g2.flatten().tolist()))
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_barrier_shard_reduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=100) for _ in range(2))
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, slow, owned, n, g1, g2 in res:
        assert slow == 2.0
        assert owned == [1.0] * 32
        assert n == 11
        assert g1 == [1.5] * 5 and g2 == [15.0] * 6


def _metrics_worker(rank, world, port, q):
    """Data-parallel metric state (dist_reduce_fx="sum" in the reference, precipitation_metrics.py:26-34) and the loader's
    rank sharding, exercised the way an N-GPU run uses them."""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    P.init_from_env(backend="gloo")
    import numpy as np
    from oracle import metrics_oracle as MO
    from smaat_unet_b200 import data as D
    from smaat_unet_b200.metrics import PrecipitationMetrics
    m = PrecipitationMetrics(device="cpu")
    st = MO.new_state()
    batches = MO.metric_batches(n_batches=4, nan_batch=99)
    for p, t in batches[rank::world]:                      # each rank folds its own batches (here through the oracle)
        MO.update(st, p, t)
    m._totals[:8] = torch.tensor([st["total_loss"], st["total_loss_denorm"], st["total_samples"], st["total_pixels"],
                                  st["total_tn"], st["total_fp"], st["total_fn"], st["total_tp"]], dtype=torch.float64)
    m.sync_across_ranks()
    out = {k: float(v) for k, v in m.compute().items()}
    ds = D.precipitation_maps_oversampled_shard(np.zeros((10, 13, 2, 2), np.float32), 12, 1)
    mine = D.PinnedBatchLoader(ds, batch_size=2, shuffle=True, seed=3, rank=rank, world=world, pin_memory=False).epoch_indices()
    q.put((rank, out, mine))
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_metric_state_sum_and_loader_sharding():
    from oracle import metrics_oracle as MO
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_metrics_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=100) for _ in range(2))
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    st = MO.new_state()
    for p, t in MO.metric_batches(n_batches=4, nan_batch=99):
        MO.update(st, p, t)
    want = MO.compute(st)
    for rank, out, mine in res:                             # every rank ends with the metrics of ALL batches
        for k, v in want.items():
            assert abs(out[k] - v) <= 1e-6 * max(abs(v), 1e-12), (rank, k)
    assert sorted(res[0][2] + res[1][2]) == list(range(10)) and len(res[0][2]) == len(res[1][2]) == 5
