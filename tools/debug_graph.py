import sys, torch
sys.path.insert(0, ".")
import smaat_unet_b200 as S
from smaat_unet_b200.train import TrainSession
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
S.set_pointwise_mode(mode)
torch.manual_seed(3)
B, S_ = 2, 32
base = S.SmaAt_UNet(12, 1, kernels_per_layer=2).cuda().train()
sd = {k: v.clone() for k, v in base.state_dict().items()}
xs = [torch.rand(B, 12, S_, S_, device="cuda") for _ in range(4)]
ys = [torch.rand(B, S_, S_, device="cuda") for _ in range(4)]
res = {}
for g in (False, True):
    m = S.SmaAt_UNet(12, 1, kernels_per_layer=2).cuda().train(); m.load_state_dict(sd)
    sess = TrainSession(m, B, (12, S_, S_), lr=1e-3, use_graph=g)
    res[g] = [float(sess.step(x, y)) for x, y in zip(xs, ys)]
    # repeat the same batch: loss must change only through the update
    res[(g, "p")] = {k: v.clone() for k, v in m.state_dict().items()}
print(mode, "eager", res[False]); print(mode, "graph", res[True])
worst = sorted(((float((res[(False, 'p')][k].float() - res[(True, 'p')][k].float()).abs().max()), k) for k in sd), reverse=True)[:6]
print(worst)
