"""Time smaat_pw1x1_fwd over (K, N, stats, mode) at P = 288^2, B = 32 with CUDA events (tools, not a test)."""
import sys, torch
sys.path.insert(0, ".")
from smaat_unet_b200 import ops
B, H = 32, 288
P = H * H
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for K, N in [(24, 64), (128, 64), (256, 64), (64, 128), (128, 128)]:
    x = torch.rand(B, K, H, H, device="cuda")
    w = torch.randn(N, K, 1, 1, device="cuda") * 0.1
    sh = torch.randn(N, device="cuda")
    split = ops.split_tf32(w.view(N, -1))
    out = torch.empty(B, N, H, H, device="cuda")
    for mode in ("tf32", "tf32x3"):
        for st in (False, True):
            stats = ops.new_stats(N, x.device) if st else None
            ms = t(lambda: ops.pw1x1(x, w, None, sh, False, mode=mode, w_split=split if mode == "tf32x3" else None, stats=stats, out=out))
            gb = 4 * B * P * (K + N) / 1e9
            print(f"K={K:4d} N={N:4d} {mode:7s} stats={int(st)}  {ms:7.3f} ms  {gb/ms*1e3:7.0f} GB/s")
