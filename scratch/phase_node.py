"""Phase cycles of the node-level forward kernel (mlp_fwd_kernel<128,2,2>; needs lib/libgp_b200_prof.so)."""
import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/graph-physics_b200")
from graphphysics_b200 import ops
from graphphysics_b200.synthetic import cylinder_flow_batch
from graphphysics_b200.models.processors import EncodeProcessDecode
dev = torch.device("cuda:0")
b = cylinder_flow_batch(32, seed=0).to(dev)
N, H = b.x.shape[0], 128
m = EncodeProcessDecode(1, 11, 3, 2, hidden_size=H).to(dev); eng = m.engine
bf = torch.bfloat16
x = torch.randn(N, H, device=dev).to(bf); agg = torch.randn(N, H, device=dev).to(bf); P = torch.randn(N, 3 * H, device=dev).to(bf)
x2 = torch.empty_like(x); h2n = torch.empty_like(x)
names = ["issue loads+gather+tmem_st", "wait loads/sync", "mma wait (x4)", "hidden epilogue (x3)", "norm epilogue", "slot sync after epi", "resid ld + segment walk", "output pass + sync", "(unused)", "mma issue L0 (+next idx)", "mma issue L1 (+L2 prefetch)", "mma issue L2 (+h2 store)", "mma issue L3"]
for it in range(3):
    prof = torch.zeros(32, dtype=torch.int64, device=dev)
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    eng._mlp(eng.node[0], N, agg, H, x2, H, resid=x, init=P, init_off0=2 * H, save_h2=h2n, prof=prof)
    en.record(); torch.cuda.synchronize()
    p = prof.cpu().tolist(); tiles = max(p[15], 1)
    print(f"iter {it}: {st.elapsed_time(en)*1e3:.0f} us, tiles {tiles}, cycles/tile per phase:")
    tot = 0
    for i, n in enumerate(names):
        print(f"    {n:32s} {p[i]/tiles:9.0f}"); tot += p[i] / tiles
    print(f"    {'total':32s} {tot:9.0f}")
