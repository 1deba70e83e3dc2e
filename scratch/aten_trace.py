"""Which ATen kernels run inside one eager training step of the benchmark configuration (and from where)."""
import sys, collections
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/graph-physics_b200")
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from graphphysics_b200.synthetic import cylinder_flow_batch
from graphphysics_b200.training.loop import Trainer
dev = torch.device("cuda:0")
tr = Trainer(bench.CONFIG if hasattr(bench, "CONFIG") else bench.config_2(), learning_rate=1e-4, num_steps=1000, warmup=10, device=dev)
b = cylinder_flow_batch(32, seed=0).to(dev)
for _ in range(3): tr.training_step(b)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    tr.training_step(b); torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_stack_n=6) if e.device_time_total > 0 and e.key.startswith("aten::")]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:40]:
    st = [s for s in e.stack if "graphphysics_b200" in s or "bench" in s][:2]
    print(f"{e.key:28s} n={e.count:3d} dev={e.device_time_total:8.1f}us  {' | '.join(x.split('graphphysics_b200/')[-1] for x in st)}")
print("total aten device us:", sum(e.device_time_total for e in rows))
