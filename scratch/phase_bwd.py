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
g = get_csr(b.edge_index, N); s = eng.edge[0]
bf = torch.bfloat16
e = torch.randn(E, H, device=dev).to(bf); h2 = torch.randn(E, H, device=dev).abs().to(bf); P = torch.randn(N, 3*H, device=dev).to(bf)
dE = torch.randn(E, H, device=dev).to(bf); dagg = torch.randn(N, H, device=dev).to(bf)
delta2 = torch.empty((E, H), dtype=bf, device=dev); dEn = torch.empty_like(delta2); d1 = torch.empty_like(delta2)
dPd = torch.empty((N, H), device=dev, dtype=bf); bnd = torch.empty(ops.seg_bnd_size(E, H, backward=True), device=dev)
names = ["P0 issue", "P0 wait+sync", "gather combine+publish", "P1 MMA", "E1", "P2 MMA", "E2 norm bwd", "P3 MMAs", "E3", "P4 issue+copyout+walk", "P4 MMA wait", "E4+output"]
sn = eng.node[0]
agg = torch.randn(N, H, device=dev).to(bf); h2n = torch.randn(N, H, device=dev).abs().to(bf); dX = torch.randn(N, H, device=dev)
delta2n = torch.empty((N, H), dtype=bf, device=dev); dagg_o = torch.empty((N, H), dtype=bf, device=dev); dQ = torch.empty((N, H), dtype=bf, device=dev)
for it in range(2):
  for stage in ("B", "A", "node B", "node A"):
    prof = torch.zeros(32, dtype=torch.int64, device=dev)
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    if stage == "B":
        ops.mlp_bwd_stage(E, H, a=h2, ka=H, wa=s.packed[2], ba=s.bias[2], wb=s.packed[3], bb=s.bias[3], partials=eng.partials,
                          norm_scale=s.scale, gy=dE, gy_gather=dagg, gy_idx=g.dst, out=delta2, mask_by_ain=True, prof=prof)
    elif stage == "node B":
        ops.mlp_bwd_stage(N, H, a=h2n, ka=H, wa=sn.packed[2], ba=sn.bias[2], wb=sn.packed[3], bb=sn.bias[3], partials=eng.partials,
                          norm_scale=sn.scale, gy=dX, out=delta2n, mask_by_ain=True, prof=prof)
    elif stage == "node A":
        ops.mlp_bwd_stage(N, H, a=agg, ka=H, wa=sn.packed[0], ba=sn.bias[0], wb=sn.packed[1], bb=sn.bias[1], partials=eng.partials,
                          delta_b=delta2n, out=dagg_o, delta_a_out=dQ, init=P, init_off0=2 * H, prof=prof)
    else:
        ops.mlp_bwd_stage(E, H, a=e, ka=H, wa=s.packed[0], ba=s.bias[0], wb=s.packed[1], bb=s.bias[1], partials=eng.partials,
                          init=P, init_off0=0, init_off1=H, idx0=g.dst, idx1=g.src, two_inits=True, delta_b=delta2, out=dEn,
                          out_resid=dE, delta_a_out=d1, seg_id=g.dst, seg_out=dPd, seg_bnd=bnd, prof=prof)
    en.record(); torch.cuda.synchronize()
    p = prof.cpu().tolist(); tiles = p[15]
    if it == 1:
        print(f"stage {stage}: {st.elapsed_time(en)*1e3:.0f} us, tiles {tiles}; cycles/tile:")
        tot = 0
        for i, n in enumerate(names):
            print(f"    {n:28s} {p[i]/tiles:8.0f}"); tot += p[i]/tiles
        print(f"    {'total':28s} {tot:8.0f}")
        print(f"    CTA 0 timeline, ns from entry: prologue done {p[16]}, tile loop done {p[17]}, partials written {p[18]}")
