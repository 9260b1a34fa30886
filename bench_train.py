#!/usr/bin/env python
"""bench_train.py -- BASELINE configs[2]/[3]: SmaAt-UNet training step (fwd + bwd + Adam) on B200.

  python bench_train.py [--batch 32] [--steps 10] [--warmup 3] [--mode tf32x3]
  torchrun --nproc-per-node N bench_train.py --global-batch 256      # configs[3]: DDP, one flat gradient all-reduce

Loss = mse_loss(pred.squeeze(1), y, reduction="sum") / B and Adam(lr=1e-3) as in the reference
(models/regression_lightning.py:47-65).  BatchNorm statistics stay per rank (no SyncBatchNorm in the reference).
Not part of the driver's bench contract; prints one JSON line for profiles/.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import smaat_unet_b200 as S  # noqa: E402
from smaat_unet_b200 import parallel as PAR  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--global-batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", default="tf32x3")
    a = ap.parse_args()
    rank, world, local = PAR.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    PAR.init_from_env("nccl", dev)
    S.set_pointwise_mode(a.mode)
    B = a.global_batch // world if a.global_batch else a.batch
    torch.manual_seed(0)
    model = S.SmaAt_UNet(12, 1, kernels_per_layer=2).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    params = [p for p in model.parameters()]
    gen = torch.Generator().manual_seed(1 + rank)
    x = torch.rand((B, 12, 288, 288), generator=gen).to(dev)
    y = torch.rand((B, 288, 288), generator=gen).to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        pred = model(x)
        loss = torch.nn.functional.mse_loss(pred.squeeze(1), y, reduction="sum") / B
        loss.backward()
        if world > 1:
            PAR.allreduce_flat_([p.grad for p in params], average=True)
        opt.step()
        return loss

    for _ in range(a.warmup):
        loss = step()
    PAR.barrier(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = S._lib.launch_count()
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    PAR.barrier(dev)
    ms = PAR.reduce_max(e0.elapsed_time(e1), dev)
    if rank == 0:
        print(json.dumps({"task": "train step fwd+bwd+Adam", "frames_per_s": world * B * a.steps / (ms * 1e-3), "ms_per_step": ms / a.steps,
                          "n_gpus": world, "batch_per_gpu": B, "pointwise": a.mode, "final_loss": float(loss),
                          "gpu_launches": int(S._lib.launch_count() - n0), "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
