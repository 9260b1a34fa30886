"""Stage timers of the fused DS-conv kernel (CTA 0) for one shape.  usage: ds_timing.py C S Cout [mode]"""
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
buf = (ctypes.c_ulonglong * 16)()
for _ in range(2): ops.dsconv(x, dw_w, dw_b, k, pw_w, sc, sh, True, mode=mode, w_split=split)
torch.cuda.synchronize(); lib.smaat_debug_dsconv_timing(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.dsconv(x, dw_w, dw_b, k, pw_w, sc, sh, True, mode=mode, w_split=split); e1.record()
torch.cuda.synchronize(); lib.smaat_debug_dsconv_timing(buf)
v = list(buf)
names = ["prod wait in", "prod wait A free", "prod compute", "prod chunks", "mma wait A", "mma wait B", "mma wait acc", "mma issue", "mma chunks", "epi wait acc", "epi work", "epi tiles", "kernel"]
print(f"C={C} S={H} Cout={Cout} {mode}: {e0.elapsed_time(e1):.3f} ms")
for n, val in zip(names, v): print(f"  {n:18s} {val:12d}")
if v[3]: print(f"  per producer chunk: wait_in {v[0]/v[3]:.0f}  wait_A {v[1]/v[3]:.0f}  compute {v[2]/v[3]:.0f} cycles")
if v[8]: print(f"  per mma chunk: wait_A {v[4]/v[8]:.0f} wait_B {v[5]/v[8]:.0f} issue {v[7]/v[8]:.0f};  per tile wait_acc {v[6]/max(v[11],1):.0f}")
if v[11]: print(f"  per epilogue tile: wait {v[9]/v[11]:.0f} work {v[10]/v[11]:.0f};  kernel cycles/tile {v[12]/v[11]:.0f}")
