"""Time the fused DS-conv kernel for the SmaAt-UNet layer shapes (CUDA events).  usage: time_ds.py [mode]"""
import sys, torch
sys.path.insert(0, ".")
from smaat_unet_b200 import ops
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
if len(sys.argv) > 2: ops.set_dsconv_impl(sys.argv[2])
B, k = 32, 2
tot = 0
for C, H, Cout in [(12, 288, 64), (64, 288, 64), (64, 144, 128), (128, 144, 128), (256, 72, 128), (256, 144, 128), (128, 144, 64), (128, 288, 64), (64, 288, 64), (128, 72, 256), (256, 72, 256), (512, 72, 256)]:
    x = torch.rand(B, C, H, H, device="cuda")
    dw_w = torch.randn(k * C, 1, 3, 3, device="cuda"); dw_b = torch.randn(k * C, device="cuda")
    pw_w = torch.randn(Cout, k * C, 1, 1, device="cuda") * 0.1
    sc = torch.rand(Cout, device="cuda") + 0.5; sh = torch.randn(Cout, device="cuda")
    split = ops.split_tf32(pw_w.view(Cout, -1))
    f = lambda: ops.dsconv(x, dw_w, dw_b, k, pw_w, sc, sh, True, mode=mode, w_split=split)
    if f() is None:
        print(f"C={C:4d} S={H:4d} N={Cout:4d}: not taken"); continue
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5; tot += ms
    gb = 4 * B * H * H * (C + Cout) / 1e9
    print(f"C={C:4d} S={H:4d} N={Cout:4d} {mode}: {ms:7.3f} ms {gb/ms*1e3:6.0f} GB/s {2*B*H*H*2*C*Cout/ms/1e9:6.1f} TF")
print("sum", round(tot, 3), "ms")
