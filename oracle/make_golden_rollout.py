"""Generate tests/golden/rollout.npz from the UNMODIFIED reference (build container only; TEST
INFRASTRUCTURE).

    python oracle/make_golden_rollout.py

An autoregressive roll-out over the reference's own mock trajectory (tests/mock_vtu/cylinder_0..5.vtu,
stored in tests/golden/cylinder_mesh.npz): the reference's Simulator / EncodeProcessDecode modules make the
predictions; the loop around them restates LightningModule._make_prediction / validation_step /
on_validation_epoch_end (graphphysics/training/lightning_module.py:27-35, 375-409, 411-492; the Lightning
class itself is not importable here), `use_previous_data` off."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gp_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ref = ref_shim.import_reference()
    processors, simulator = ref["processors"], ref["simulator"]
    NT = ref["nodetype"].NodeType
    from torch_geometric.data import Data
    torch.set_num_threads(4)
    z = np.load(f"{OUT}/cylinder_mesh.npz")
    pos = z["points"][:, :2].astype(np.float32)
    ei = O.face_to_edge(z["triangles"].astype(np.int64), len(pos))
    ea = O.edge_features(pos, ei)
    vel = z["velocity"]                                   # [6, N, 2]
    N = len(pos)
    nt = np.zeros(N, np.int64)                            # geometric node types (same rule as make_golden.py)
    nt[pos[:, 0] < pos[:, 0].min() + 1e-6] = int(NT.INFLOW)
    nt[pos[:, 0] > pos[:, 0].max() - 1e-6] = int(NT.OUTFLOW)
    nt[(pos[:, 1] < pos[:, 1].min() + 1e-6) | (pos[:, 1] > pos[:, 1].max() - 1e-6)] = int(NT.WALL_BOUNDARY)
    ei_t, ea_t, pos_t = torch.from_numpy(ei), torch.from_numpy(ea), torch.from_numpy(pos)
    T = vel.shape[0] - 1
    frames = [torch.cat([torch.from_numpy(vel[t]), torch.from_numpy(nt)[:, None].float(), torch.full((N, 1), float(t))], 1)
              for t in range(T)]
    ys = [torch.from_numpy(vel[t + 1]) for t in range(T)]

    torch.manual_seed(11)
    index = dict(feature_index_start=0, feature_index_end=2, output_index_start=0, output_index_end=2, node_type_index=2)
    model = processors.EncodeProcessDecode(3, 2 + 9, 3, 2, hidden_size=64)
    sim = simulator.Simulator(node_input_size=11, edge_input_size=3, output_size=2, model=model, device=torch.device("cpu"), **index)
    # a short fit so that the roll-out is not that of a random network: normaliser statistics + 40 Adam steps
    opt = torch.optim.AdamW(sim.parameters(), lr=2e-3)
    lossf = ref["loss"].L2Loss()
    sim.train()
    for it in range(40):
        t = it % T
        b = Data(x=frames[t].clone(), y=ys[t], pos=pos_t, edge_index=ei_t, edge_attr=ea_t)
        net, tgt, _ = sim(b)
        loss = lossf(tgt, net, b.x[:, 2], masks=[NT.NORMAL, NT.OUTFLOW])
        opt.zero_grad()
        loss.backward()
        opt.step()
    sim.eval()

    # lightning_module.py:375-409 (_make_prediction), 27-35 (build_mask), 446-451 and 466-486 (the two RMSEs)
    last, preds = None, []
    for t in range(T):
        b = Data(x=frames[t].clone(), y=ys[t], pos=pos_t, edge_index=ei_t, edge_attr=ea_t)
        if last is not None:
            b.x[:, 0:2] = last.detach()
        node_type = b.x[:, 2]
        mask = torch.logical_not(torch.logical_or(node_type == NT.NORMAL, node_type == NT.OUTFLOW))
        with torch.no_grad():
            _, _, predicted = sim(b)
        predicted[mask] = b.y[mask]
        last = predicted
        preds.append(predicted.clone())
    P, Y = torch.cat(preds), torch.cat(ys)
    rmse_1 = torch.sqrt(((preds[0] - ys[0]) ** 2).mean()).item()
    rmse_all = torch.sqrt(((P - Y) ** 2).mean()).item()
    print(f"roll-out over {T} frames, N={N}: val_1step_rmse {rmse_1:.6f}  val_all_rollout_rmse {rmse_all:.6f}  (final fit loss {loss.item():.4f})")
    np.savez_compressed(f"{OUT}/rollout.npz", pos=pos, edge_index=ei, edge_attr=ea, frames=torch.stack(frames).numpy(),
                        ys=torch.stack(ys).numpy(), predictions=torch.stack(preds).numpy(),
                        val_1step_rmse=np.float64(rmse_1), val_all_rollout_rmse=np.float64(rmse_all),
                        **{"sd/" + k: v.detach().numpy() for k, v in sim.state_dict().items()})


if __name__ == "__main__":
    main()
