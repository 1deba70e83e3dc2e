import sys, torch
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
x = torch.randn(N, H, device=dev).to(torch.bfloat16); e = torch.randn(E, H, device=dev).to(torch.bfloat16)
P = torch.randn(N, 3*H, device=dev).to(torch.bfloat16)
bnd = torch.empty(ops.seg_bnd_size(E, H), device=dev); agg = torch.empty((N, H), device=dev, dtype=torch.bfloat16); e2 = torch.empty_like(e)
names = ["issue loads+gather+tmem_st", "wait loads/sync", "mma wait (x4)", "hidden epilogue (x3)", "norm epilogue", "slot sync after epi", "resid ld + segment walk", "output pass + sync", "(unused)", "mma issue L0 (+next idx)", "mma issue L1 (+L2 prefetch)", "mma issue L2 (+h2 store)", "mma issue L3"]
import os
h2 = torch.empty((E, H), device=dev, dtype=torch.bfloat16) if os.environ.get('SAVE_H2') else None
for it in range(3):
    prof = torch.zeros(16, dtype=torch.int64, device=dev)
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    eng._mlp(eng.edge[0], E, e, H, e2, H, resid=e, init=P, init_off0=0, init_off1=H, idx0=g.dst, idx1=g.src, two_inits=True,
             seg_id=g.dst, seg_out=agg, seg_bnd=bnd, prof=prof, save_h2=h2)
    en.record(); torch.cuda.synchronize()
    p = prof.cpu().tolist(); tiles = p[15]
    print(f"iter {it}: {st.elapsed_time(en)*1e3:.0f} us, tiles {tiles}, cycles/tile per phase:")
    tot = 0
    for i, n in enumerate(names):
        print(f"    {n:32s} {p[i]/tiles:9.0f}"); tot += p[i]/tiles
    print(f"    {'total':32s} {tot:9.0f}")
