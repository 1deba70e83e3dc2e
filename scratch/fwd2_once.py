"""One launch sequence of the CTA-pair edge kernel on the benchmark batch (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/graph-physics_b200")
from graphphysics_b200 import ops
from graphphysics_b200.synthetic import cylinder_flow_batch
from graphphysics_b200.graph import get_csr
from graphphysics_b200.models.processors import EncodeProcessDecode
dev = torch.device("cuda:0")
b = cylinder_flow_batch(32, seed=0).to(dev)
N, E, H = b.x.shape[0], b.edge_index.shape[1], 128
m = EncodeProcessDecode(1, 11, 3, 2, hidden_size=H).to(dev); eng = m.engine
g = get_csr(b.edge_index, N)
e = torch.randn(E, H, device=dev).to(torch.bfloat16)
P = torch.randn(N, 3*H, device=dev).to(torch.bfloat16)
bnd = torch.empty(ops.seg_bnd_size(E, H), device=dev); agg = torch.empty((N, H), device=dev, dtype=torch.bfloat16); e2 = torch.empty_like(e)
h2 = torch.empty((E, H), device=dev, dtype=torch.bfloat16)
for it in range(int(os.environ.get("ITERS", "3"))):
    eng._mlp(eng.edge[0], E, e, H, e2, H, resid=e, init=P, init_off0=0, init_off1=H, idx0=g.dst, idx1=g.src, two_inits=True,
             seg_id=g.dst, seg_out=agg, seg_bnd=bnd, save_h2=h2)
torch.cuda.synchronize()
