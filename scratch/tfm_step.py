"""Two eager training steps of the coarse-aneurysm transformer (config 4) -- target of the ncu launch list."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "graph-physics_b200"))
import numpy as np, torch
from graphphysics_b200 import preprocessing as P
from graphphysics_b200.graph import Data
from graphphysics_b200.training.loop import Trainer
dev = torch.device("cuda:0")
gold = os.path.join(ROOT, "tests", "golden")
cfg = json.load(open(os.path.join(gold, "training_configs.json")))["coarse-aneurysm"]
a = np.load(os.path.join(gold, "aneurysm_mesh.npz"))
n = a["points"].shape[0]
ei = P.cells_to_edge_index(torch.from_numpy(a["tets"].T.astype(np.int64)).to(dev), n)
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(n, 15, device=dev, generator=gen); x[:, 14] = 0
b = Data(x=x, y=torch.randn(n, 3, device=dev, generator=gen), edge_index=ei)
tr = Trainer(cfg, learning_rate=1e-4, num_steps=1000, warmup=10, device=dev, seed=0)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(steps):
    tr.training_step(b)
torch.cuda.synchronize()
print("loss", float(tr._loss[0]))
