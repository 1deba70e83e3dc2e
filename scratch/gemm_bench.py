"""Warm CUDA-event timing of the gp_gemm shapes of the coarse-aneurysm Transformer block."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "graph-physics_b200"))
import torch
from graphphysics_b200 import dense
dev = torch.device("cuda:0")
R, H = 22535, 64
x16 = torch.randn(R, H, device=dev).to(torch.bfloat16)
x32 = torch.randn(R, H, device=dev)
g16 = torch.randn(R, 3 * H, device=dev).to(torch.bfloat16)
w = torch.randn(H, H, device=dev); b = torch.randn(H, device=dev)
w1 = torch.randn(3 * H, H, device=dev); b1 = torch.randn(3 * H, device=dev)
w3 = torch.randn(H, 3 * H, device=dev)
dy = torch.randn(R, H, device=dev); dy3 = torch.randn(R, 3 * H, device=dev)
def t(name, fn, n=50):
    """GPU time per call: `n` calls captured into one CUDA graph (no host launch cost between them), replayed 5 times."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    if n < 10:
        return
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / (5 * n) * 1e3:8.1f} us")
only = sys.argv[1] if len(sys.argv) > 1 else None
cases = {
 "fwd bf16->bf16 N=64": lambda: dense.lin_fwd(x16, w, b, out_bf16=True),
 "fwd bf16->f32 N=64 +resid": lambda: dense.lin_fwd(x16, w, b, resid=x32),
 "fwd bf16->f32 N=192": lambda: dense.lin_fwd(x16, w1, b1),
 "fwd K=192 N=64 +resid": lambda: dense.lin_fwd(g16, w3, b, resid=x32),
 "dgrad N=64": lambda: dense.lin_dgrad(dy, w),
 "dgrad K=192 -> 64": lambda: dense.lin_dgrad(dy3, w1),
 "dgrad accumulate": lambda: dense.lin_dgrad(dy, w, out=x32),
 "wgrad 64x64 +bias": lambda: dense.lin_wgrad(dy, x16, bias=True),
 "wgrad 192x64 +bias": lambda: dense.lin_wgrad(dy3, x16, bias=True),
 "wgrad 64x192 +bias": lambda: dense.lin_wgrad(dy, g16, bias=True),
 "rmsnorm fwd": lambda: dense.norm_fwd(x32, b),
 "rmsnorm bwd": lambda: dense.norm_bwd(x32, b, None, dy),
 "gelu fwd": lambda: dense.gelu_fwd(dy3, dy3),
 "empty alloc only": lambda: torch.empty((R, H), device=dev),
}
for k, f in cases.items():
    if only is None or only in k:
        t(k, f, n=50 if only is None else 3)
