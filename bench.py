#!/usr/bin/env python
"""bench.py -- SmaAt-UNet forward frames/sec on B200 (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode tf32x3|tf32|fp32]

* own arm (``--impl b200``): N ranks (torchrun for N>1), each with the full model and its own
  shard of B=32 synthetic 12x288x288 frames per step (weak scaling, eval forward has no
  collective -- SURVEY 8e).  ``value`` = frames/s with inputs resident in HBM (CUDA-graph replay,
  CUDA events, barrier + synchronize both sides, max over ranks); ``e2e`` = the same through
  ``InferenceSession.submit/collect`` with pinned HOST buffers (H2D + D2H inside the timed region);
  ``roofline`` = the depthwise kernel (the metric's named kernel) timed live with CUDA events,
  algorithmic bytes / time vs MEASURED_PEAKS.json; ``cpu_baseline`` = the oracle's torch CPU port
  on a bounded sample (rank 0, N=1 only).
* reference arm (``--impl reference``): the reference's CPU path (oracle/torch_port.py: same
  ATen/oneDNN kernels as the reference modules; /root/reference is not on the GPU box) on all
  host threads, each step a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
B_PER_GPU, C_IN, SIZE = 32, 12, 288


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def randomise_bn(model, gen):
    """make eval-mode BN non-trivial."""
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) + 0.5)
                m.weight.copy_(torch.rand(m.weight.shape, generator=gen) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=gen) * 0.1)


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons DURING the timed region (NVML, 20 Hz)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def usable_cpus():
    """Host threads this process can really use: affinity mask, capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per))))
    except Exception:
        pass
    return n


def cpu_port_time(n_frames, reps, threads):
    """Oracle leg (allowed to import oracle/): reference algorithm on the host cores."""
    import numpy as np
    from oracle import torch_port as TP
    from oracle.cases import cast_sd, fill_schema, smaat_unet_schema
    torch.set_num_threads(threads)
    sd = TP.to_torch_sd(cast_sd(fill_schema(smaat_unet_schema(C_IN, 1, 2), 0), np.float32))
    x = torch.rand(n_frames, C_IN, SIZE, SIZE)
    with torch.no_grad():
        TP.smaat_unet_forward(x[:1], sd)          # warm-up (oneDNN primitive creation)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            TP.smaat_unet_forward(x, sd)
            ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args, rank):
    if rank != 0:
        return
    threads = usable_cpus()
    n = 4
    ts = cpu_port_time(n, args.warmup + args.steps, threads)[args.warmup:]
    sec = sum(ts)
    fps = n * len(ts) / sec
    sample = f"{len(ts)} steps x {n} frames of 12x{SIZE}x{SIZE} (B=32 workload, bounded), torch CPU fp32, {threads} threads"
    out = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / len(ts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: full SmaAt-UNet forward, batch=32, 12->1ch 288x288 (bounded sample of 4 frames/step)",
                   "kernels_per_layer": 2, "eval": True},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default=os.environ.get("SMAAT_PW_MODE", "tf32x3"), choices=["tf32x3", "tf32", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the reported-only tf32 measurement")
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference", "timing rules: W >= 3"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    import smaat_unet_b200 as S
    from smaat_unet_b200 import parallel as PAR
    from smaat_unet_b200.engine import InferenceSession

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    PAR.init_from_env("nccl", dev)
    S.set_pointwise_mode(args.mode)

    gen = torch.Generator().manual_seed(0)
    torch.manual_seed(0)
    model = S.SmaAt_UNet(C_IN, 1, kernels_per_layer=2)
    randomise_bn(model, gen)
    model = model.to(dev).eval()
    sess = InferenceSession(model, B_PER_GPU, (C_IN, SIZE, SIZE), device=dev, use_graph=not args.no_graph)

    # two resident input batches (alternated); a step touches ~40 GB of activations >> 126 MB L2
    xs = [torch.rand((B_PER_GPU, C_IN, SIZE, SIZE), generator=gen).to(dev) for _ in range(2)]
    host = [torch.rand((B_PER_GPU, C_IN, SIZE, SIZE), generator=gen).pin_memory() for _ in range(2)]

    def barrier():
        PAR.barrier(dev)

    def reduce_max(v):
        return PAR.reduce_max(v, dev)

    # ---------------- device-resident throughput ("value") ----------------
    for i in range(args.warmup):
        sess.forward(xs[i % 2])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = S._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        sess.forward(xs[i % 2])
    e1.record()
    barrier()
    ms = reduce_max(e0.elapsed_time(e1))
    eager_launches = S._lib.launch_count() - n0
    clocks = sampler.result()
    fps = world * B_PER_GPU * args.steps / (ms * 1e-3)
    launches = (sess.launches_per_forward * args.steps) if sess.graph is not None else eager_launches

    # ---------------- end to end through the public API, host buffers ----------------
    for i in range(args.warmup):
        sess.submit(host[i % 2])
        sess.collect()
    barrier()
    t0 = time.perf_counter()
    chk = 0.0
    for i in range(args.steps):
        sess.submit(host[i % 2])
        if i >= 1:
            chk += float(sess.collect()[0, 0, 0, 0])       # read the result on the host
    chk += float(sess.collect()[0, 0, 0, 0])
    torch.cuda.synchronize()
    e2e_s = reduce_max(time.perf_counter() - t0)
    barrier()
    e2e_fps = world * B_PER_GPU * args.steps / e2e_s

    # ---------------- reported-only: same measurement in the single-pass TF32 mode ----------------
    # (what the reference itself computes on a GPU: cuDNN allow_tf32=True; ~1e-3 relative error instead of 1e-6)
    alt = None
    if args.mode == "tf32x3" and not args.no_alt:
        S.set_pointwise_mode("tf32")
        sess2 = InferenceSession(model, B_PER_GPU, (C_IN, SIZE, SIZE), device=dev, use_graph=not args.no_graph)
        for i in range(args.warmup):
            sess2.forward(xs[i % 2])
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(args.steps):
            sess2.forward(xs[i % 2])
        a1.record()
        barrier()
        ams = reduce_max(a0.elapsed_time(a1))
        alt = {"pointwise": "tf32", "value": world * B_PER_GPU * args.steps / (ams * 1e-3), "unit": "frames/s", "ms_per_step": ams / args.steps}
        del sess2
        S.set_pointwise_mode(args.mode)

    # ---------------- roofline: per-kernel timing, CUDA events on the launching stream ----------------
    roof, roof_dw, kernels = None, None, {}
    if rank == 0:
        hbm, src = peaks()
        with torch.no_grad():
            model(xs[0])
            torch.cuda.synchronize()
            with S.ops.profile() as prof:
                for i in range(3):
                    model(xs[i % 2])
            agg = prof.summary()
            # the last DS conv carries the fused OutConv epilogue (its own ABI entry): same kernel, count it with the others
            oc = agg.pop("smaat_dsconv_outconv_fwd", None)
            if oc is not None and "smaat_dsconv_fwd" in agg:
                for k_ in ("launches", "ms", "bytes", "flops"):
                    agg["smaat_dsconv_fwd"][k_] += oc[k_]
            elif oc is not None:
                agg["smaat_dsconv_fwd"] = oc
            if os.environ.get("SMAAT_BENCH_LAYERS"):
                for name, a in prof.summary(by_shape=True).items():
                    if "[" in name:
                        print(f"# {name:40s} {a['ms'] / 3:8.3f} ms  {a['bytes'] / a['ms'] / 1e6:7.0f} GB/s  {a['flops'] / a['ms'] / 1e9:7.1f} TF", file=sys.stderr)
        for name, a in agg.items():
            gbs = a["bytes"] / (a["ms"] * 1e-3) / 1e9 if a["ms"] > 0 else 0.0
            kernels[name] = {"launches_per_step": a["launches"] // 3, "ms_per_step": a["ms"] / 3, "algorithmic_GB_per_step": a["bytes"] / 3e9,
                             "achieved_GBps": gbs, "frac_hbm": gbs / hbm, "tflops": a["flops"] / (a["ms"] * 1e-3) / 1e12 if a["ms"] > 0 else 0.0}
        # dominant kernel of the measured path (by time): its own algorithmic bytes / its own time
        KNAMES = {"smaat_dsconv_fwd": "dsconv_fused_kernel (depthwise 3x3 -> tcgen05 pointwise -> BN/ReLU, one kernel)",
                  "smaat_pw1x1_fwd": "pw1x1_tc_kernel", "smaat_dw3x3_fwd": "dw3x3_kernel"}
        dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        d = kernels[dom]
        # DRAM traffic per launch from the ncu --set full capture of the same kernels (profiles/), GB; None if not captured
        roof = {"kernel": KNAMES.get(dom, dom) + f" ({d['launches_per_step']} launches/step)", "bound": "hbm",
                "achieved": d["achieved_GBps"], "peak": hbm, "unit": "GB/s", "frac": d["frac_hbm"], "peak_source": src,
                "traffic": 9.928e8 if dom == "smaat_dsconv_fwd" else None,
                "traffic_note": "ncu --set full, dram read+write of ONE launch (up3.0: C256->128 @144^2): 992.8 MB vs 1019 MB "
                                "algorithmic for that launch (profiles/r01d_ncu_summary.md); the 9 launches of a step differ in size"
                if dom == "smaat_dsconv_fwd" else "",
                "algorithmic_bytes_per_step": d["algorithmic_GB_per_step"] * 1e9, "ms_per_step": d["ms_per_step"],
                "note": "fused DS conv: depthwise producers, tcgen05 issue and epilogue are balanced within ~10% (stage timers in "
                        "profiles/r01_dsconv_stage_timers.txt); its algorithmic bytes are 2.8x fewer than dw+pw unfused (DESIGN.md 5)"
                if dom == "smaat_dsconv_fwd" else ""}
        # the metric's named kernel -- "depthwise % HBM roofline": the standalone depthwise kernel over ALL 18 layers
        # (fusion switched off for this measurement pass only)
        S.set_fused_dsconv(False)
        with torch.no_grad():
            model(xs[0])
            torch.cuda.synchronize()
            with S.ops.profile() as prof2:
                for i in range(3):
                    model(xs[i % 2])
            a2 = prof2.summary().get("smaat_dw3x3_fwd")
        S.set_fused_dsconv(True)
        if a2:
            g2 = a2["bytes"] / (a2["ms"] * 1e-3) / 1e9
            roof_dw = {"kernel": f"dw3x3_kernel ({a2['launches'] // 3} launches/step, all DS layers, unfused pass)", "bound": "hbm",
                       "achieved": g2, "peak": hbm, "unit": "GB/s", "frac": g2 / hbm, "peak_source": src,
                       "traffic": 4.02e9, "traffic_note": "ncu dram read+write for the largest launch (up4.0): 4.02 GB vs 4.08 GB algorithmic (profiles/r01_ncu_summary_v1.md)",
                       "algorithmic_bytes_per_step": a2["bytes"] / 3, "ms_per_step": a2["ms"] / 3}

    # ---------------- CPU baseline (oracle port), rank 0, N=1 only ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = usable_cpus()
        n = 8
        ts = cpu_port_time(n, 1, threads)
        cpu = {"value": n / ts[0], "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": f"1 timed forward of {n} frames 12x{SIZE}x{SIZE} after a 1-frame warm-up; oracle/torch_port.py (torch CPU fp32, {threads} threads)"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: full SmaAt-UNet forward (eval), batch=32 per GPU, 12->1ch 288x288, kernels_per_layer=2",
                       "global_batch": B_PER_GPU * world, "pointwise": args.mode, "cuda_graph": sess.graph is not None,
                       "parallelism": f"batch-sharded x{world}, no collective",
                       "l2": "inputs alternate between 2 buffers; a step streams ~40 GB of activations (>> 126 MB L2)"},
            "roofline": roof, "depthwise_roofline": roof_dw, "kernels": kernels, "cpu_baseline": cpu,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": sess.h2d_bytes_per_step,
                    "d2h_bytes_per_step": sess.d2h_bytes_per_step, "ms_per_step": 1e3 * e2e_s / args.steps, "checksum": chk},
            "alt_mode": alt, "clocks": clocks, "gpu_launches": int(launches),
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
