"""Stage timers of the TMEM-operand fused DS-conv kernel (CTA 0) for one shape.  usage: dt_timing.py C S Cout [mode]"""
import os, sys, ctypes, torch
os.environ["SMAAT_DSCONV_TIMING"] = "1"
sys.path.insert(0, ".")
from smaat_unet_b200 import ops, _lib
C, H, Cout = (int(a) for a in sys.argv[1:4]); mode = sys.argv[4] if len(sys.argv) > 4 else "tf32x3"
B, k = 32, 2
x = torch.rand(B, C, H, H, device="cuda")
dw_w = torch.randn(k * C, 1, 3, 3, device="cuda"); dw_b = torch.randn(k * C, device="cuda")
pw_w = torch.randn(Cout, k * C, 1, 1, device="cuda") * 0.1
sc = torch.rand(Cout, device="cuda") + 0.5; sh = torch.randn(Cout, device="cuda")
split = ops.split_tf32(pw_w.view(Cout, -1))
lib = _lib.load()
ops.set_dsconv_impl("tmem")
buf = (ctypes.c_ulonglong * 24)()
for _ in range(2): ops.dsconv(x, dw_w, dw_b, k, pw_w, sc, sh, True, mode=mode, w_split=split)
torch.cuda.synchronize(); lib.smaat_debug_dsconv_tmem_timing(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.dsconv(x, dw_w, dw_b, k, pw_w, sc, sh, True, mode=mode, w_split=split); e1.record()
torch.cuda.synchronize(); lib.smaat_debug_dsconv_tmem_timing(buf)
v = list(buf)
print(f"C={C} S={H} Cout={Cout} {mode}: {e0.elapsed_time(e1):.3f} ms (timers on)   kernel cycles {v[12]}")
d = lambda a, b: a / max(b, 1)
print(f"  producer g0 ({v[3]} units): wait_in {d(v[0],v[3]):.0f}  wait_A_free {d(v[1],v[3]):.0f}  compute+st {d(v[2],v[3]):.0f}")
print(f"  mma ({v[8]} units): wait_A {d(v[4],v[8]):.0f}  wait_B {d(v[5],v[8]):.0f}  wait_acc {d(v[6],v[8]):.0f}  issue {d(v[7],v[8]):.0f}")
print(f"  epilogue ({v[11]} pairs): wait {d(v[9],v[11]):.0f}  work {d(v[10],v[11]):.0f}   kernel cycles/pair {d(v[12],v[11]):.0f}  /unit {d(v[12],v[8]):.0f}")
print(f"  tma ({v[14]} units): wait_free {d(v[13],v[14]):.0f};  stager ({v[17]}): wait_free {d(v[15],v[17]):.0f} total {d(v[16],v[17]):.0f};  wloader ({v[19]}): wait_free {d(v[18],v[19]):.0f}")
n = 148
cb = (ctypes.c_ulonglong * (3 * n))()
lib.smaat_debug_dsconv_tmem_cta_timing(cb, n)
c = list(cb)
t0 = min(c[3 * i] for i in range(n) if c[3 * i])
rows = sorted(((c[3 * i + 1] - t0) / 1e3, (c[3 * i] - t0) / 1e3, c[3 * i + 2], i) for i in range(n) if c[3 * i])
ends = [r[0] for r in rows]; durs = sorted(r[0] - r[1] for r in rows)
print(f"  per-CTA ({len(rows)} CTAs): start spread {max(r[1] for r in rows):.1f} us; end min/median/max {ends[0]:.1f}/{ends[len(ends)//2]:.1f}/{ends[-1]:.1f} us; "
      f"duration min/median/max {durs[0]:.1f}/{durs[len(durs)//2]:.1f}/{durs[-1]:.1f} us")
print("  slowest 8 (end us, sm, cta):", [(round(r[0], 1), r[2], r[3]) for r in rows[-8:]], " fastest 4:", [(round(r[0], 1), r[2], r[3]) for r in rows[:4]])
