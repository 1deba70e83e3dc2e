"""Generate tests/golden/* from the UNMODIFIED reference (build container only; TEST INFRASTRUCTURE).

    python oracle/make_golden.py

Imports /root/reference through oracle/ref_shim.py, runs the reference's own modules on seeded
inputs and stores inputs, weights (state_dict), outputs, losses and gradients.  The committed
fixtures pin oracle/gp_oracle.py (tests/test_oracle_golden.py) and, through the oracle, the CUDA
path.  Also converts the two mesh fixtures of the reference's tests to compact .npz files."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gp_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.vtu_reader import read_vtu  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF = ref_shim.REFERENCE_ROOT


def sd_np(sd):
    return {"sd/" + k: v.detach().cpu().numpy() for k, v in sd.items()}


def node_types(pos, rng):
    t = np.zeros(len(pos), np.int64)
    t[pos[:, 0] < 1e-6] = 4
    t[pos[:, 0] > 1.6 - 1e-6] = 5
    t[(pos[:, 1] < 1e-6) | (pos[:, 1] > 0.41 - 1e-6)] = 6
    return t


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.import_reference()
    layers, processors, simulator, loss_m, sched_m = (ref[k] for k in ("layers", "processors", "simulator", "loss", "scheduler"))
    from torch_geometric.data import Data
    torch.set_num_threads(4)

    # ---- mesh fixtures of the reference's tests -> compact npz
    cyl = [read_vtu(f"{REF}/tests/mock_vtu/cylinder_{i}.vtu") for i in range(6)]
    vel = np.stack([np.stack([c["point_data"]["velocity_x"], c["point_data"]["velocity_y"]], -1) for c in cyl]).astype(np.float32)
    np.savez_compressed(f"{OUT}/cylinder_mesh.npz", points=cyl[0]["points"].astype(np.float32),
                        triangles=cyl[0]["cells"].astype(np.int32), velocity=vel)
    an = read_vtu(f"{REF}/tests/mock_vtu_aneurysm/aneurysm_0.vtu")
    np.savez_compressed(f"{OUT}/aneurysm_mesh.npz", points=an["points"].astype(np.float32), tets=an["cells"].astype(np.int32))

    # ---- small synthetic graph shared by the model goldens
    pos, tris = O.grid_tri_mesh(14, 9, jitter=0.3, seed=3, hole=(0.4, 0.2, 0.08))
    ei = O.face_to_edge(tris, len(pos))
    ea = O.edge_features(pos, ei)
    N, E = len(pos), ei.shape[1]
    rng = np.random.default_rng(0)
    ei_t, ea_t = torch.from_numpy(ei), torch.from_numpy(ea)

    # 1) EncodeProcessDecode: forward, loss-free scalar, gradients
    for name, (L, H) in {"epd_l2_h32": (2, 32), "epd_l2_h64": (2, 64)}.items():
        torch.manual_seed(1)
        model = processors.EncodeProcessDecode(L, 11, 3, 2, hidden_size=H)
        x = torch.randn(N, 11)
        G = torch.randn(N, 2)
        out = model(Data(x=x, edge_index=ei_t, edge_attr=ea_t))
        (out * G).sum().backward()
        grads = {"grad/" + k: p.grad.numpy() for k, p in model.named_parameters()}
        np.savez_compressed(f"{OUT}/{name}.npz", x=x.numpy(), edge_attr=ea, edge_index=ei, G=G.numpy(), out=out.detach().numpy(),
                            L=L, H=H, **sd_np(model.state_dict()), **grads)

    # 2) GraphNetBlock alone
    torch.manual_seed(2)
    blk = layers.GraphNetBlock(32)
    x, e = torch.randn(N, 32), torch.randn(E, 32)
    ox, oe = blk(x, ei_t, e)
    np.savez_compressed(f"{OUT}/graphnet_block_h32.npz", x=x.numpy(), e=e.numpy(), edge_index=ei, out_x=ox.detach().numpy(),
                        out_e=oe.detach().numpy(), **sd_np(blk.state_dict()))

    # 3) EncodeTransformDecode on the DGL branch (coarse-aneurysm.json shape, 2 layers)
    torch.manual_seed(3)
    tm = processors.EncodeTransformDecode(2, 23, 3, hidden_size=64, num_heads=4)
    x = torch.randn(N, 23)
    G = torch.randn(N, 3)
    out = tm(Data(x=x, edge_index=ei_t))
    (out * G).sum().backward()
    grads = {"grad/" + k: p.grad.numpy() for k, p in tm.named_parameters()}
    np.savez_compressed(f"{OUT}/transformer_l2_h64.npz", x=x.numpy(), edge_index=ei, G=G.numpy(), out=out.detach().numpy(),
                        **sd_np(tm.state_dict()), **grads)

    # 4) Simulator + L2Loss + optimizer: three training steps and one eval step (cylinder.json layout)
    torch.manual_seed(4)
    index = dict(feature_index_start=0, feature_index_end=2, output_index_start=0, output_index_end=2, node_type_index=2)
    model = processors.EncodeProcessDecode(2, 2 + 9, 3, 2, hidden_size=32)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    sim = simulator.Simulator(node_input_size=11, edge_input_size=3, output_size=2, model=model, device=torch.device("cpu"), **index)
    nt = node_types(pos, rng)
    frames = []
    for f in range(4):
        v = torch.from_numpy(rng.standard_normal((N, 2)).astype(np.float32))
        frames.append(torch.cat([v, torch.from_numpy(nt)[:, None].float(), torch.full((N, 1), float(f))], 1))
    ys = [fr[:, :2] + 0.1 * torch.from_numpy(rng.standard_normal((N, 2)).astype(np.float32)) for fr in frames]
    opt = torch.optim.AdamW(sim.parameters(), lr=1e-3, weight_decay=1e-4, betas=(0.9, 0.95))
    sch = sched_m.CosineWarmupScheduler(opt, warmup=2, max_iters=10)
    lossf = loss_m.L2Loss()
    NT = ref["nodetype"].NodeType
    losses, lrs = [], []
    sim.train()
    for s in range(3):
        b = Data(x=frames[s], y=ys[s], pos=torch.from_numpy(pos), edge_index=ei_t, edge_attr=ea_t)
        net, tgt, _ = sim(b)
        loss = lossf(tgt, net, b.x[:, 2], masks=[NT.NORMAL, NT.OUTFLOW])
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(sim.parameters(), 1.0)
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
        losses.append(loss.item())
    sim.eval()
    with torch.no_grad():
        b = Data(x=frames[3], y=ys[3], pos=torch.from_numpy(pos), edge_index=ei_t, edge_attr=ea_t)
        net, tgt, outp = sim(b)
    np.savez_compressed(f"{OUT}/train_steps.npz", pos=pos, edge_index=ei, edge_attr=ea, frames=torch.stack(frames).numpy(),
                        ys=torch.stack(ys).numpy(), losses=np.array(losses), lrs=np.array(lrs), eval_net=net.numpy(),
                        eval_target=tgt.numpy(), eval_outputs=outp.numpy(),
                        **{"sd0/" + k: v.numpy() for k, v in sd0.items()},
                        **{"sd3/" + k: v.detach().numpy() for k, v in sim.state_dict().items()})

    # 5) small pieces with pinned behaviour in the reference's own tests
    torch.manual_seed(5)
    rn = layers.RMSNorm(16)
    xx = torch.randn(7, 16)
    nz = layers.Normalizer(5, device="cpu")
    d1, d2 = torch.randn(9, 5), torch.randn(4, 5)
    n1 = nz(d1)
    n2 = nz(d2)
    np.savez_compressed(f"{OUT}/small_ops.npz", rms_x=xx.numpy(), rms_out=rn(xx).detach().numpy(), norm_d1=d1.numpy(),
                        norm_d2=d2.numpy(), norm_n1=n1.numpy(), norm_n2=n2.numpy(), norm_inv=nz.inverse(n2).numpy(),
                        sched=np.array([sched_m.CosineWarmupScheduler.get_lr_factor(sch, e) for e in range(12)]))
    # 6) the `model` / `index` / `training` sections of the shipped training configs, verbatim
    import json
    cfgs = {n: {k: json.load(open(f"{REF}/training_config/{n}.json"))[k] for k in ("model", "index", "training")}
            for n in ("cylinder", "plate", "coarse-aneurysm")}
    json.dump(cfgs, open(f"{OUT}/training_configs.json", "w"), indent=1)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f"  {f:32s} {os.path.getsize(os.path.join(OUT, f)) / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
