"""Large-mesh run (BASELINE config 5 shape: ~1M nodes, ~7M undirected / 14M directed edges, 15 MP layers, H=128).
    python scratch/mesh1m.py [n]                       one GPU, unpartitioned forward+backward
    torchrun --nproc-per-node P scratch/mesh1m.py [n]  node-partitioned over P GPUs (halo exchange fwd + bwd)
Prints ms per forward+backward (max over ranks) and edges/s per MP layer."""
import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/graph-physics_b200")
import numpy as np, torch
import torch.distributed as dist
from graphphysics_b200.graph import Data, get_csr
from graphphysics_b200.models.processors import EncodeProcessDecode
from graphphysics_b200.synthetic import kuhn_box_graph, mesh_edge_attr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))); torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
t0 = time.time()
pos, ei = kuhn_box_graph(n, n, n)
ea = mesh_edge_attr(pos, ei)
N, E, L, H = len(pos), ei.shape[1], 15, 128
torch.manual_seed(0)
model = EncodeProcessDecode(L, 12, 4, 3, hidden_size=H).to(dev)
gen = torch.Generator().manual_seed(1)
x = torch.randn(N, 12, generator=gen); G = torch.randn(N, 3, generator=gen)
if rank == 0: print(f"mesh {n}^3: N={N} E={E} directed, built in {time.time()-t0:.1f}s", flush=True)

def timed(fn, steps=3):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)

if world == 1:
    xd, Gd = x.to(dev), G.to(dev); eid, ead = torch.from_numpy(ei).to(dev), torch.from_numpy(ea).to(dev)
    eng = model.engine; g = get_csr(eid, N)
    def step():
        out, _, ctx = eng.forward(xd, ead, g, save=True)
        eng.backward(ctx, Gd)
    ms = timed(step)
else:
    from graphphysics_b200.dist.partition import build_local_graphs, partition_nodes
    from graphphysics_b200.dist.partitioned import PartitionedEPD
    t0 = time.time()
    owner = partition_nodes(pos, world)
    lg = build_local_graphs(ei, owner, world)[rank]
    part = PartitionedEPD(model, lg, world, dist.group.WORLD)
    halo = int(sum(len(v) for v in lg.recv.values()))
    print(f"rank {rank}: owned {lg.num_owned} ghosts {halo} local edges {lg.edge_index_local.shape[1]} (partition {time.time()-t0:.1f}s)", flush=True)
    xd, ead = x.to(dev), torch.from_numpy(ea).to(dev)
    Gown = G[torch.from_numpy(lg.owned)].to(dev)
    def step():
        out, ctx = part.forward(xd, ead, save=True)
        part.backward(ctx, Gown)
    ms = timed(step)
if rank == 0:
    print(f"world {world}: {ms:.1f} ms per forward+backward, {E * L / (ms * 1e-3):.3e} edges/s per MP layer, "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
if world > 1:
    dist.barrier(); torch.cuda.synchronize(); os._exit(0)
