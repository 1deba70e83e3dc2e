"""Helper of tests/test_mlp_bwd_gpu.py::test_specialised_backward_instantiations_equal_general (not a test
module): runs both stages of the edge-MLP backward on fixed inputs and saves every output.  The library
picks the compile-time specialised instantiations unless GP_BWD_GENERAL is set (read once per process), so
the test runs this file in two processes:  python tests/cmp_bwd_modes.py <outdir>"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "graph-physics_b200"))
OUTDIR = sys.argv[1] if len(sys.argv) > 1 else "/tmp"
from graphphysics_b200 import ops
from graphphysics_b200.synthetic import cylinder_flow_batch
from graphphysics_b200.graph import get_csr
from graphphysics_b200.models.processors import EncodeProcessDecode
dev = torch.device("cuda:0"); torch.manual_seed(0)
b = cylinder_flow_batch(2, seed=0).to(dev)
N, E, H = b.x.shape[0], b.edge_index.shape[1], 128
m = EncodeProcessDecode(1, 11, 3, 2, hidden_size=H).to(dev); eng = m.engine
g = get_csr(b.edge_index, N); s = eng.edge[0]; bf = torch.bfloat16
e = torch.randn(E, H, device=dev).to(bf); h2 = torch.randn(E, H, device=dev).abs().to(bf); P = torch.randn(N, 3*H, device=dev).to(bf)
dE = torch.randn(E, H, device=dev).to(bf); dagg = torch.randn(N, H, device=dev).to(bf)
delta2 = torch.empty((E, H), dtype=bf, device=dev); dEn = torch.empty_like(delta2); d1 = torch.empty_like(delta2)
dPd = torch.empty((N, H), device=dev, dtype=bf); bnd = torch.empty(ops.seg_bnd_size(E, H, backward=True), device=dev)
part = eng.partials_all[:eng._region_elems]
out = {}
gB = ops.mlp_bwd_stage(E, H, a=h2, ka=H, wa=s.packed[2], ba=s.bias[2], wb=s.packed[3], bb=s.bias[3], partials=part,
                       norm_scale=s.scale, gy=dE, gy_gather=dagg, gy_idx=g.dst, out=delta2, mask_by_ain=True)
stride = ops.bwd_layout(H, H, H)[5]
out["B_part"] = part[: gB * stride].clone().view(gB, stride).sum(0); out["delta2"] = delta2.clone()
gA = ops.mlp_bwd_stage(E, H, a=e, ka=H, wa=s.packed[0], ba=s.bias[0], wb=s.packed[1], bb=s.bias[1], partials=part,
                       init=P, init_off0=0, init_off1=H, idx0=g.dst, idx1=g.src, two_inits=True, delta_b=delta2, out=dEn,
                       out_resid=dE, delta_a_out=d1, seg_id=g.dst, seg_out=dPd, seg_bnd=bnd)
ops.seg_fixup(g.rowptr_dst, H, bnd, dPd, backward=True)
out["A_part"] = part[: gA * stride].clone().view(gA, stride).sum(0); out["dEn"] = dEn.clone(); out["d1"] = d1.clone(); out["dPd"] = dPd.clone()
torch.cuda.synchronize()
tag = "general" if os.environ.get("GP_BWD_GENERAL") else "fast"
torch.save({k: v.float().cpu() for k, v in out.items()}, os.path.join(OUTDIR, f"cmp_{tag}.pt"))
if tag == "fast":
    ref = torch.load(os.path.join(OUTDIR, "cmp_general.pt"))
    o = ops.bwd_layout(H, H, H)
    names = ["dWb", "dWa", "dbb", "dba", "dsc"]
    for k, v in out.items():
        v = v.float().cpu(); r = ref[k]
        if k.endswith("part"):
            for i, nme in enumerate(names):
                lo, hi = o[i], (o[i + 1] if i + 1 < 5 else o[5])
                d = (v[lo:hi] - r[lo:hi]).abs().max().item(); print(f"{k}:{nme:4s} maxdiff {d:.3e}  ref max {r[lo:hi].abs().max().item():.3e} nan={bool(torch.isnan(v[lo:hi]).any())}")
        else:
            print(f"{k:8s} maxdiff {(v - r).abs().max().item():.3e} nan={bool(torch.isnan(v).any())}")
if tag == "fast":
    v = out["dEn"].float().cpu(); r = ref["dEn"]
    bad = torch.isnan(v) | ((v - r).abs() > 1e-3)
    rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
    print("general has nan:", bool(torch.isnan(r).any()), "bad rows", rows.numel(), rows[:10].tolist(), rows[-5:].tolist(), "bad cols", cols.numel(), cols[:8].tolist(), cols[-4:].tolist())
    print("tile-local rows of first bad rows:", (rows[:20] % 128).tolist())
