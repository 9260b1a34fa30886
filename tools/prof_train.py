import sys, torch
sys.path.insert(0, ".")
import smaat_unet_b200 as S
from smaat_unet_b200 import ops
B = 32
torch.manual_seed(0)
model = S.SmaAt_UNet(12, 1, kernels_per_layer=2).cuda().train()
x = torch.rand(B, 12, 288, 288, device="cuda"); y = torch.rand(B, 288, 288, device="cuda")
def step():
    model.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(model(x).squeeze(1), y, reduction="sum") / B
    loss.backward()
for _ in range(2): step()
torch.cuda.synchronize()
with ops.profile() as prof:
    step()
agg = prof.summary()
tot = sum(a["ms"] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k:30s} n={a['launches']:4d} {a['ms']:8.2f} ms {100*a['ms']/tot:5.1f}%  {a['bytes']/max(a['ms'],1e-9)/1e6:7.0f} GB/s {a['flops']/max(a['ms'],1e-9)/1e9:7.1f} TF")
print("total kernel ms", tot)
print("--- GEMMs by shape")
bs = prof.summary(by_shape=True)
for k, a in sorted(bs.items(), key=lambda kv: -kv[1]["ms"]):
    if "pw1x1" in k:
        print(f"{k:48s} n={a['launches']:3d} {a['ms']:7.3f} ms {a['bytes']/max(a['ms'],1e-9)/1e6:7.0f} GB/s {a['flops']/max(a['ms'],1e-9)/1e9:7.1f} TF")
