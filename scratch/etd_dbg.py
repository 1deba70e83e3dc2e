import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'graph-physics_b200')
import torch
from oracle import gp_oracle as O
from graphphysics_b200.graph import Data
from graphphysics_b200.models.processors import EncodeTransformDecode
from tests.util import l2_rel
from tests.test_dense_gpu import _mesh_graph
dev = torch.device("cuda:0")
torch.manual_seed(5)
n, ei = _mesh_graph(16, seed=2)
m = EncodeTransformDecode(3, 23, 3, hidden_size=64, num_heads=4)
sd = {k: v.detach().clone().double().requires_grad_(True) for k, v in m.state_dict().items()}
x, dy = torch.randn(n, 23), torch.randn(n, 3)
ref = O.etd_forward(sd, x.double(), ei, 3, 4, mode="bf16")
(ref * dy.double()).sum().backward()
m = m.to(dev)
out = m(Data(x=x.to(dev), edge_index=ei.to(dev)))
(out * dy.to(dev)).sum().backward()
print("out", l2_rel(out, ref))
for name, p in m.named_parameters():
    r = sd[name].grad
    print(f"{name:50s} {float((p.grad.double().cpu() - r).norm() / r.norm()):.2e}  |r|={float(r.norm()):.2e}")
