"""Event trace of CTA 0 of the TMEM-operand fused DS-conv kernel (SMAAT_DSCONV_TIMING=2).  usage: dt_trace.py C S Cout [first_unit n_units]"""
import os, sys, ctypes, torch
os.environ["SMAAT_DSCONV_TIMING"] = "2"
sys.path.insert(0, ".")
from smaat_unet_b200 import ops, _lib
C, H, Cout = (int(a) for a in sys.argv[1:4])
u0 = int(sys.argv[4]) if len(sys.argv) > 4 else 64
nu = int(sys.argv[5]) if len(sys.argv) > 5 else 12
B, k = 32, 2
x = torch.rand(B, C, H, H, device="cuda")
dw_w = torch.randn(k * C, 1, 3, 3, device="cuda"); dw_b = torch.randn(k * C, device="cuda")
pw_w = torch.randn(Cout, k * C, 1, 1, device="cuda") * 0.1
sc = torch.rand(Cout, device="cuda") + 0.5; sh = torch.randn(Cout, device="cuda")
split = ops.split_tf32(pw_w.view(Cout, -1))
lib = _lib.load()
ops.set_dsconv_impl("tmem")
for _ in range(3): ops.dsconv(x, dw_w, dw_b, k, pw_w, sc, sh, True, mode="tf32x3", w_split=split)
torch.cuda.synchronize()
N = 256
buf = (ctypes.c_longlong * (16 * N))()
lib.smaat_debug_dsconv_tmem_trace(buf, N)
t = [list(buf[16 * u:16 * u + 16]) for u in range(N)]
nch = (C + 15) // 16
base = t[u0][0]
names = ["tma", "p:in", "p:cmp", "p:Afree", "p:Ast", "m:ops", "m:h0>", "m:h0<", "m:h1>", "m:h1<", "e:h0>", "e:h0<", "e:h1>", "e:h1<"]
print(f"C={C} S={H} Cout={Cout}: {nch} units per pair; cycles relative to unit {u0}'s TMA issue; flags={os.environ.get('SMAAT_DT_FLAGS', '0')}")
print("unit  " + " ".join(f"{n:>8s}" for n in names))
for u in range(u0, min(u0 + nu, N)):
    row = [f"{t[u][k] - base:8d}" if t[u][k] else "       ." for k in range(14)]
    if u % nch != 0:
        row[10:14] = ["       ."] * 4
    print(f"{u:4d}  " + " ".join(row))
# steady-state averages over units u0 .. N-1
def avg(f):
    v = [f(u) for u in range(u0, N - 1) if all(t[u][k] for k in range(10)) and all(t[u + 1][k] for k in range(10))]
    return sum(v) / max(len(v), 1)
print(f"avg per unit: period {avg(lambda u: t[u + 1][9] - t[u][9]):.0f} | stencil {avg(lambda u: t[u][2] - t[u][1]):.0f}  wait A free {avg(lambda u: t[u][3] - t[u][2]):.0f}  A store {avg(lambda u: t[u][4] - t[u][3]):.0f}"
      f" | A stored -> issuer sees {avg(lambda u: t[u][5] - t[u][4]):.0f}  sees -> h0 start {avg(lambda u: t[u][6] - t[u][5]):.0f}  h0 issue {avg(lambda u: t[u][7] - t[u][6]):.0f}  gap {avg(lambda u: t[u][8] - t[u][7]):.0f}  h1 issue {avg(lambda u: t[u][9] - t[u][8]):.0f}"
      f" | h1 issued(u) -> h0 start(u+1) {avg(lambda u: t[u + 1][6] - t[u][9]):.0f} | h1 issued(u) -> A stage free seen by producer of u+AS: see rows")
