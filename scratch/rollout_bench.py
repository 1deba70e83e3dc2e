"""Roll-out throughput on one CylinderFlow-shaped mesh (C1 topology size, 15 MP layers, H=128): eager vs CUDA-graph."""
import sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/graph-physics_b200")
from graphphysics_b200.synthetic import cylinder_flow_batch
from graphphysics_b200.training.loop import Trainer
dev = torch.device("cuda:0")
cfg = {"model": {"type": "epd", "message_passing_num": 15, "hidden_size": 128, "node_input_size": 2, "output_size": 2, "edge_input_size": 3},
       "index": {"feature_index_start": 0, "feature_index_end": 2, "output_index_start": 0, "output_index_end": 2, "node_type_index": 2}}
tr = Trainer(cfg, learning_rate=1e-3, num_steps=10, warmup=2, device=dev, seed=0)
frames = [cylinder_flow_batch(1, seed=s).to(dev) for s in range(1)] * 200
for graphed in (False, True):
    tr.enable_cuda_graph(graphed)
    tr.rollout(frames[:5]); torch.cuda.synchronize()
    t0 = time.time(); r = tr.rollout(frames); torch.cuda.synchronize(); dt = time.time() - t0
    print(f"{'graph' if graphed else 'eager'}: {len(frames) / dt:.0f} roll-out steps/s ({dt / len(frames) * 1e3:.3f} ms per frame), N={frames[0].x.shape[0]} E={frames[0].edge_index.shape[1]}")
