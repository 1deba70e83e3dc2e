import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'graph-physics_b200')
import torch
from oracle import gp_oracle as O
from graphphysics_b200 import dense
from graphphysics_b200.models.layers import build_mlp
from tests.util import l2_rel
dev = torch.device("cuda:0")
for (i, h, o, ln) in [(23, 64, 64, True), (64, 64, 3, False), (64, 64, 64, True), (24, 64, 8, False)]:
    torch.manual_seed(1)
    m = build_mlp(i, h, o, layer_norm=ln)
    sd = {"m." + k: v.detach().clone().double().requires_grad_(True) for k, v in m.state_dict().items()}
    R = 500
    x, dy = torch.randn(R, i), torch.randn(R, o)
    x64 = x.double().requires_grad_(True)
    ref = O.dense_mlp(x64, sd, "m", layer_norm=ln, mode="bf16")
    (ref * dy.double()).sum().backward()
    m = m.to(dev)
    xd = x.to(dev).requires_grad_(True)
    out = dense.mlp4(m, xd)
    (out * dy.to(dev)).sum().backward()
    print(i, h, o, ln, "out", l2_rel(out, ref), "dx", l2_rel(xd.grad, x64.grad))
    for name, p in m.named_parameters():
        r = sd["m." + name].grad
        print(f"   {name:20s} {float((p.grad.double().cpu() - r).norm() / r.norm()):.2e}")
