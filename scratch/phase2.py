"""Phase cycles of the CTA-pair edge kernel (needs a library built with -DGP_FWD2_PROF: GP_EXTRA_FLAGS=-DGP_FWD2_PROF
bash build.sh).  Counters are taken by one thread per role of CTA 0, slot 0."""
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
x = torch.randn(N, H, device=dev).to(torch.bfloat16); e = torch.randn(E, H, device=dev).to(torch.bfloat16)
P = torch.randn(N, 3*H, device=dev).to(torch.bfloat16)
bnd = torch.empty(ops.seg_bnd_size(E, H), device=dev); agg = torch.empty((N, H), device=dev, dtype=torch.bfloat16); e2 = torch.empty_like(e)
names = ["epi: bookkeeping + requests", "epi: wait staged rows", "epi: acc pre-load", "epi: wait e tile", "epi: wait MMA (x4)", "epi: hidden epilogue (x3)",
         "epi: norm sumsq+exchange", "epi: wait drain", "epi: norm scale+write u", "drain: resid requests", "drain: wait u", "drain: e'=e+u", "drain: segment walk",
         "prod: issue+idx", "prod: wait staging"]
h2 = torch.empty((E, H), device=dev, dtype=torch.bfloat16) if os.environ.get('SAVE_H2') else None
for it in range(3):
    prof = torch.zeros(32, dtype=torch.int64, device=dev)
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    eng._mlp(eng.edge[0], E, e, H, e2, H, resid=e, init=P, init_off0=0, init_off1=H, idx0=g.dst, idx1=g.src, two_inits=True,
             seg_id=g.dst, seg_out=agg, seg_bnd=bnd, prof=prof, save_h2=h2)
    en.record(); torch.cuda.synchronize()
    p = prof.cpu().tolist(); tiles = max(p[15], 1)
    print(f"iter {it}: {st.elapsed_time(en)*1e3:.0f} us, tiles of CTA0/slot0 {tiles}, cycles/tile per phase:")
    for i, n in enumerate(names):
        print(f"    {n:32s} {p[i]/tiles:9.0f}")
    for i, n in ((16, "epi: l=2 wait h2 store read"), (17, "epi: hidden convert (x3)"), (18, "epi: fences (x3)"), (19, "epi: l=1 sync + store issue"),
                 (20, "pre: requests + wait accumulator"), (21, "pre: wait staged rows"), (22, "pre: sums + TMEM stores")):
        print(f"    {n:32s} {p[i]/tiles:9.0f}")
    print(f"    timeline of CTA 0 (ns from entry): loop start {p[24]}, loop end {p[25]}, exit {p[26]}")
    print(f"    {'epilogue total':32s} {(sum(p[:9])+sum(p[16:20]))/tiles:9.0f}   drain total {sum(p[9:13])/tiles:9.0f}   producer total {sum(p[13:15])/tiles:9.0f}")
