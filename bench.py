#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native graph-physics hot path.

Workload (BASELINE.json configs[1]): CylinderFlow-shaped synthetic meshes (~1.9k nodes, ~11k
directed edges each), EncodeProcessDecode with 15 message-passing layers, hidden 128, batch 32
graphs per GPU, data-parallel over --gpus ranks (weak scaling).  One step = one full training
step (Simulator forward, masked L2, backward, gradient all-reduce, clip, AdamW, LR schedule).

Metric: edges/s per message-passing layer (fwd+bwd) = directed edges in the global batch x MP
layers / step time.  `value` is measured with the batch resident in HBM; `e2e` goes through the
public Trainer.training_step with the batch in pinned host memory (H2D copy every step, loss
read back every step).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference      # the reference's own modules (oracle/_ref) on the host CPU cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "graph-physics_b200"))

import torch  # noqa: E402

CONFIG = {
    "model": {"type": "epd", "message_passing_num": 15, "hidden_size": 128, "node_input_size": 2, "output_size": 2,
              "edge_input_size": 3},
    "index": {"feature_index_start": 0, "feature_index_end": 2, "output_index_start": 0, "output_index_end": 2,
              "node_type_index": 2},
}
BATCH = 32
METRIC = "edges/sec per MP layer (fwd+bwd)"
UNIT = "edges/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._stop_evt = index, [], set(), threading.Event()
        self.sm_max = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [s.strip() for s in out.split(",")]
                self.samples.append(float(f[0]))
                self.sm_max = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def run_reference(args, sample_graphs: int = BATCH, budget_s: float = 150.0, max_steps: int = 3):
    """The UNMODIFIED reference (its own EncodeProcessDecode / Simulator / L2Loss / CosineWarmupScheduler, staged by
    oracle/build_ref.py into oracle/_ref and imported through the torch_geometric / dgl stand-ins of oracle/ref_shim.py)
    running the same training step on the host CPU cores, fp32, all threads.  `sample_graphs` of the 32 graphs of one
    rank's batch per step (32 = the whole batch: same config as the GPU arm); at most `max_steps` timed steps and never
    more than `budget_s` seconds after the warm-up."""
    from graphphysics_b200.synthetic import cylinder_flow_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = CONFIG["model"]
    kind = "reference"
    try:
        from oracle.ref_trainer import ReferenceTrainer
        tr = ReferenceTrainer(CONFIG, lr=1e-4, num_steps=100000, warmup=1000, seed=0)
        step = lambda b: tr.training_step(b.x, b.y, b.pos, b.edge_index, b.edge_attr)
    except Exception as exc:                      # staged copy missing: fall back to the oracle port and SAY so
        print(f"bench: reference modules unavailable ({exc}); timing the oracle port instead", file=sys.stderr)
        from oracle.cpu_train import CpuTrainer, default_state_dict
        kind = "port"
        idx = CONFIG["index"]
        sd = default_state_dict(m["message_passing_num"], m["node_input_size"] + 9, m["edge_input_size"], m["output_size"],
                                m["hidden_size"])
        ct = CpuTrainer(sd, m["message_passing_num"], idx, m["output_size"], m["node_input_size"] + 9, m["edge_input_size"],
                        lr=1e-4, num_steps=100000, warmup=1000)
        step = lambda b: ct.training_step(b.x, b.y, b.edge_attr, b.edge_index)
    b = cylinder_flow_batch(sample_graphs, seed=0)
    E = b.edge_index.shape[1]
    warm = max(args.warmup_ref, 1)
    small = cylinder_flow_batch(1, seed=1)
    step(small)                                   # thread pool / allocator warm-up on one graph
    for _ in range(warm):
        step(b)
    t0 = time.perf_counter()
    done = 0
    while done < max_steps and (done < 1 or time.perf_counter() - t0 < budget_s):
        step(b)
        done += 1
    dt = (time.perf_counter() - t0) / done
    value = E * m["message_passing_num"] / dt
    return {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "steps": done, "warmup": warm,
            "sample": f"{sample_graphs} of the {BATCH} graphs of one batch ({b.x.shape[0]} nodes, {E} directed edges), "
                      f"full train step, fp32, {done} steps after {warm} warm-up, {cores} threads",
            "ms_per_step": dt * 1e3}


def secondary_configs(dev, steps: int = 10, warmup: int = 3):
    """The other BASELINE.json configurations that fit one GPU, each as a full training step through Trainer with
    CUDA-graph replay (outside the headline's timed regions; `--no-secondary` skips them):
      configs[0]  cylinder.json verbatim (epd, 5 layers, hidden 32) on the reference's mock cylinder mesh
      configs[2]  DeformingPlate-shaped sample with world edges (graph built on the device), epd 15 x 128, edge_in 4
      configs[3]  coarse-aneurysm.json (transformer, 10 blocks, hidden 64, 4 heads) on the reference's mock aneurysm mesh
    The attention kernel gets its own roofline line: E * (4H + 8) algorithmic bytes per forward launch."""
    import numpy as np
    from graphphysics_b200 import ops
    from graphphysics_b200 import preprocessing as P
    from graphphysics_b200.graph import Data
    from graphphysics_b200.synthetic import deforming_plate_sample
    from graphphysics_b200.training.loop import Trainer
    hbm_peak, _, peak_src = peaks()
    gold = os.path.join(ROOT, "tests", "golden")
    cfgs = json.load(open(os.path.join(gold, "training_configs.json")))
    out = {}

    def ev_time(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def run(name, cfg, batch, extra=None, tags=()):
        tr = Trainer(cfg, learning_rate=1e-4, num_steps=100000, warmup=1000, device=dev, seed=0)
        prof = None
        if tags:                                     # per-kernel device time: eager steps, CUDA events on the launching stream
            for _ in range(2):
                tr.training_step(batch)
            ops.PROFILE.reset(tags=tags)
            ops.COUNTERS["launches"] = 0
            for _ in range(3):
                tr.training_step(batch)
            prof = ops.PROFILE.summary()
            ops.PROFILE.reset(tags=())
            launches = ops.COUNTERS["launches"] // 3
        tr.enable_cuda_graph(True)
        for _ in range(warmup):
            tr.training_step(batch)
        ms = ev_time(lambda: tr.training_step(batch), steps)
        E, L = int(batch.edge_index.shape[1]), cfg["model"]["message_passing_num"]
        rec = {"ms_per_step": ms, "train_steps_per_s": 1e3 / ms, "edges_per_s_per_layer": E * L / (ms * 1e-3),
               "nodes": int(batch.x.shape[0]), "directed_edges": E, "model": cfg["model"]["type"], "layers": L,
               "hidden": cfg["model"]["hidden_size"], "launch": "cuda-graph replay", "loss": float(tr._loss[0])}
        if prof:
            rec["kernels_ms_per_step"] = {k: v["total_ms"] / 3 for k, v in prof.items()}
            rec["gpu_launches"] = launches
        if extra:
            rec.update(extra)
        out[name] = rec
        return tr, prof

    try:        # ---- configs[0]: cylinder.json
        c = np.load(os.path.join(gold, "cylinder_mesh.npz"))
        pos = torch.from_numpy(np.ascontiguousarray(c["points"][:, :2])).to(dev)
        g = P.face_to_edge(Data(x=pos, face=torch.from_numpy(c["triangles"].T.astype(np.int64)).to(dev)))
        ea = P.edge_features(pos, g.edge_index)
        vel = torch.from_numpy(c["velocity"]).to(dev)
        n = pos.shape[0]
        x = torch.cat([vel[0], torch.zeros(n, 2, device=dev)], 1)
        b = Data(x=x, y=vel[1].contiguous(), pos=pos, edge_index=g.edge_index, edge_attr=ea)
        run("cylinder.json (configs[0])", cfgs["cylinder"], b)
    except Exception as exc:
        out["cylinder.json (configs[0])"] = {"error": str(exc)[:300]}
    try:        # ---- configs[2]: plate with world edges, graph construction on the device
        pos, tets, x_raw, y = deforming_plate_sample(seed=0)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        raw = Data(x=to(x_raw), y=to(y), pos=to(pos), tetra=to(tets.T.astype(np.int64)))
        pipe = P.build_preprocessing(world_pos_parameters={"world_pos_index_start": 0, "world_pos_index_end": 3, "node_type_index": 6})
        g = pipe(raw.clone())
        ms_pre = ev_time(lambda: pipe(raw.clone()), 10)
        cfg = {"model": {"type": "epd", "message_passing_num": 15, "hidden_size": 128, "node_input_size": 6, "output_size": 3, "edge_input_size": 4},
               "index": {"feature_index_start": 0, "feature_index_end": 6, "output_index_start": 0, "output_index_end": 3, "node_type_index": 6}}
        run("DeformingPlate-shaped, epd 15x128 with world edges (configs[2])", cfg, g,
            extra={"graph_construction_ms": ms_pre,
                   "graph_construction": "FaceToEdge + world-edge radius search + coalesce + edge features on the device, per sample"})
    except Exception as exc:
        out["DeformingPlate-shaped (configs[2])"] = {"error": str(exc)[:300]}
    try:        # ---- configs[3]: coarse-aneurysm transformer
        a = np.load(os.path.join(gold, "aneurysm_mesh.npz"))
        n = a["points"].shape[0]
        ei = P.cells_to_edge_index(torch.from_numpy(a["tets"].T.astype(np.int64)).to(dev), n)
        gen = torch.Generator(device=dev).manual_seed(0)
        x = torch.randn(n, 15, device=dev, generator=gen)
        x[:, 14] = 0.0
        b = Data(x=x, y=torch.randn(n, 3, device=dev, generator=gen), pos=torch.from_numpy(a["points"]).to(dev), edge_index=ei)
        cfg = cfgs["coarse-aneurysm"]
        _, prof = run("coarse-aneurysm.json (configs[3])", cfg, b, tags=("attn_fwd", "attn_bwd", "gemm"))
        E, H, L = int(ei.shape[1]), cfg["model"]["hidden_size"], cfg["model"]["message_passing_num"]
        if prof and "attn_fwd" in prof:
            avg_ms = prof["attn_fwd"]["total_ms"] / prof["attn_fwd"]["calls"]
            alg = E * (4 * H + 8)
            ach = alg / (avg_ms * 1e-3) / 1e9
            out["coarse-aneurysm.json (configs[3])"]["attention_roofline"] = {
                "bound": "hbm", "kernel": "csr_attention_fwd (bf16 q/k/v)", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "algorithmic_bytes": alg, "avg_launch_ms": avg_ms, "peak_source": peak_src,
                "note": "E*(4H+8) bytes per launch (SURVEY 8d); the 23 MB of k / v rows live in L2, so HBM is not what bounds this launch"}
    except Exception as exc:
        out["coarse-aneurysm.json (configs[3])"] = {"error": str(exc)[:300]}
    return out


def partition_block(dev, pg, world: int, rank: int, side: int = 72, steps: int = 3, warmup: int = 2):
    """Strong scaling of the FULL training step on one large mesh, node-partitioned with halo exchange (SURVEY §8e.2,
    BASELINE.json configs[4] scaled to what one GPU holds): a side^3 Kuhn-triangulated box (14 neighbours per interior
    node), EPD 15 x 128.  Every rank first times the unpartitioned step on the whole mesh by itself (the 1-GPU
    reference, same build, same box), then the ranks split the mesh (recursive coordinate bisection, an edge lives
    with its receiver, ghost rows exchanged per message-passing step by gp_halo_* kernels + NCCL all-to-all-v) and run
    the same step captured in one CUDA graph.  Device time, max over ranks."""
    import gc
    import numpy as np
    import torch.distributed as dist
    from graphphysics_b200.graph import Data
    from graphphysics_b200.synthetic import kuhn_box_graph, mesh_edge_attr
    from graphphysics_b200.training.loop import Trainer
    cfg = {"model": dict(CONFIG["model"], node_input_size=3, output_size=3, edge_input_size=4),
           "index": {"feature_index_start": 0, "feature_index_end": 3, "output_index_start": 0, "output_index_end": 3, "node_type_index": 3}}
    pos, ei = kuhn_box_graph(side, side, side)
    n, E = pos.shape[0], ei.shape[1]
    rng = np.random.default_rng(0)
    vel = rng.standard_normal((n, 3)).astype(np.float32)
    x = np.concatenate([vel, np.zeros((n, 2), np.float32)], 1)
    batch = Data(x=torch.from_numpy(x), y=torch.from_numpy(vel + 0.1 * rng.standard_normal((n, 3)).astype(np.float32)),
                 pos=torch.from_numpy(pos), edge_index=torch.from_numpy(ei), edge_attr=torch.from_numpy(mesh_edge_attr(pos, ei))).to(dev)

    def timed(tr, k):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            tr.training_step(batch)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / k], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    tr1 = Trainer(cfg, learning_rate=1e-4, num_steps=100000, warmup=1000, device=dev, seed=0)
    tr1.enable_cuda_graph(True)
    for _ in range(warmup):
        tr1.training_step(batch)
    ms1 = timed(tr1, steps)
    mem1 = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    loss1 = float(tr1._loss[0])
    del tr1
    from graphphysics_b200 import graph as _g
    _g._PERSISTENT.clear(); _g._CACHE.clear()
    gc.collect()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats(dev)
    trp = Trainer(cfg, learning_rate=1e-4, num_steps=100000, warmup=1000, device=dev, process_group=pg, seed=0)
    trp.enable_node_partition(batch.pos, batch.edge_index)
    trp.enable_cuda_graph(True)
    for _ in range(warmup):
        trp.training_step(batch)
    msn = timed(trp, steps)
    lg = trp._part.lg
    stats = torch.tensor([lg.num_owned, len(lg.ghosts), len(lg.edge_ids)], device=dev, dtype=torch.float32)
    mx = stats.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return {"mesh": f"{side}^3 Kuhn box, {n} nodes, {E} directed edges, EPD 15 x 128, full training step", "scaling": "strong",
            "ms_per_step_1gpu": ms1, "ms_per_step": msn, "n_gpus": world, "speedup": ms1 / msn, "efficiency": ms1 / (world * msn),
            "edges_per_s_per_layer": E * CONFIG["model"]["message_passing_num"] / (msn * 1e-3),
            "max_owned_nodes": int(mx[0]), "max_ghost_nodes": int(mx[1]), "max_local_edges": int(mx[2]),
            "gib_1gpu": mem1, "gib_per_gpu": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
            "loss_1gpu": loss1, "loss": float(trp._loss[0]), "launch": "cuda-graph replay, halo exchange inside the graph",
            "timing": "CUDA events, barrier + synchronize on both sides, max over ranks"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--steps-ref", dest="steps_ref", type=int, default=None)
    ap.add_argument("--warmup-ref", dest="warmup_ref", type=int, default=1)
    ap.add_argument("--ref-budget-s", dest="ref_budget_s", type=float, default=150.0,
                    help="--impl reference: stop timing after this many seconds (at least one step is always timed)")
    ap.add_argument("--ref-graphs", dest="ref_graphs", type=int, default=BATCH,
                    help="--impl reference: graphs per step (default: the whole batch; smaller only for smoke tests)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-partition", action="store_true", help="--gpus N > 1: skip the node-partition strong-scaling block")
    ap.add_argument("--partition-side", dest="partition_side", type=int, default=0,
                    help="--gpus N > 1: side of the box mesh of the node-partition block (side^3 nodes; 0 = by job size)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary single-GPU configurations (configs[0], [2], [3])")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"CylinderFlow-shape synthetic mesh x{BATCH} graphs per GPU, EPD 15 MP layers, hidden 128 "
                f"(BASELINE.json configs[1])")

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the whole 32-graph batch per step (the GPU arm's config); as many of the requested steps as fit the budget
        cb = run_reference(args, sample_graphs=args.ref_graphs, budget_s=args.ref_budget_s,
                           max_steps=args.steps_ref if args.steps_ref is not None else args.steps)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": cb["steps"], "warmup": cb["warmup"], "ms_per_step": cb["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "sample": cb["sample"]},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD

    from graphphysics_b200 import ops
    from graphphysics_b200.synthetic import cylinder_flow_batch
    from graphphysics_b200.training.loop import Trainer

    host = cylinder_flow_batch(BATCH, seed=rank, pin=True)
    N, E = host.x.shape[0], host.edge_index.shape[1]
    L, H = CONFIG["model"]["message_passing_num"], CONFIG["model"]["hidden_size"]
    tr = Trainer(CONFIG, learning_rate=1e-4, num_steps=100000, warmup=1000, device=dev, process_group=pg, seed=0)
    graphed = not args.no_graph
    tr.enable_cuda_graph(graphed)
    resident = host.to(dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        tr.training_step(resident)

    loss_host = []

    def step_e2e():
        # every step: one pinned-host -> device copy of a full batch and one loss read-back.  The copy is the
        # NEXT step's batch, started on a side stream right after this step's launch (double buffering,
        # as a pin_memory DataLoader does), so it overlaps the kernels instead of preceding them.
        loss = tr.training_step(None)
        tr.stage(host)
        loss_host.append(loss.item())

    tr.stage(host)

    for _ in range(args.warmup):
        step_e2e()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # kernel launches per step and per-kernel device time: one eager, instrumented step outside the
    # timed regions (under graph replay the Python wrappers do not run, the launch sequence is identical)
    tr.enable_cuda_graph(False)
    ops.PROFILE.reset(tags=("edge_fwd", "edge_bwd_B", "edge_bwd_A"))
    ops.COUNTERS["launches"] = 0
    n_prof = 3

    def step_profiled():
        # Let the host run ahead of the device: a ~25 ms spin kernel first, then the whole step is enqueued while the GPU
        # spins, so the CUDA events around a kernel bracket its DEVICE time, not the gap in which the host was still
        # building the next launch (an eager step is host-bound in places: tensor maps, argument structs).
        torch.cuda._sleep(50_000_000)
        tr.training_step(resident)

    ms_prof = timed(step_profiled, n_prof)
    launches = ops.COUNTERS["launches"] // n_prof
    prof = ops.PROFILE.summary()
    ops.PROFILE.reset(tags=())
    tr.enable_cuda_graph(graphed)
    for _ in range(2):
        step_resident()
    ms_res = timed(step_resident, args.steps)
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop()

    total_edges = E * world            # every rank holds a batch of the same shape
    value = total_edges * L * args.steps / (ms_res * 1e-3)
    e2e_value = total_edges * L * args.steps / (ms_e2e * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in (host.x, host.y, host.pos, host.edge_index, host.edge_attr))

    # roofline of the dominant kernel (SURVEY §8d algorithmic bytes, bf16 storage)
    hbm_peak, tf_peak, peak_src = peaks()
    alg_bytes = {"edge_fwd": E * (8 * H + 8) + 2 * N * H, "edge_bwd_B": E * (6 * H) + 2 * N * H,
                 "edge_bwd_A": E * (10 * H + 8) + 2 * N * H}
    alg_flops = {"edge_fwd": E * 8 * H * H, "edge_bwd_B": E * 12 * H * H, "edge_bwd_A": E * 10 * H * H}
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of the same kernels on this workload, from the
    # `ncu --set full` captures summarised in profiles/r02_final_summary.md (at or below the algorithmic bytes: the
    # gathered rows and the residual tile hit L2)
    ncu_dram_bytes = {"edge_fwd": 299.3e6, "edge_bwd_B": 284.3e6, "edge_bwd_A": 520.4e6} if (E, N, H) == (372752, 64424, 128) else {}
    roof = None
    if prof:
        top = max(prof, key=lambda k: prof[k]["total_ms"])
        avg_ms = prof[top]["total_ms"] / max(prof[top]["calls"], 1)
        ach = alg_bytes[top] / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": ncu_dram_bytes.get(top), "traffic_source": "profiles/r02_final_summary.md (ncu --set full, per launch, bytes)",
                "algorithmic_bytes": alg_bytes[top], "peak_source": peak_src, "avg_launch_ms": avg_ms,
                "share_of_step": (prof[top]["total_ms"] / n_prof) / (ms_res / args.steps),
                "tensor": {"achieved_tflops": alg_flops[top] / (avg_ms * 1e-3) / 1e12, "peak_tflops": tf_peak,
                           "frac": alg_flops[top] / (avg_ms * 1e-3) / 1e12 / tf_peak},
                "kernels_ms_per_step": {k: v["total_ms"] / n_prof for k, v in prof.items()},
                "kernels_frac": {k: alg_bytes[k] / (v["total_ms"] / max(v["calls"], 1) * 1e-3) / 1e9 / hbm_peak for k, v in prof.items()},
                "measured_in": f"{n_prof} eager steps, CUDA events around the three edge kernels on their stream; the host is kept ahead of "
                               f"the device (spin kernel before each step) so the events bracket device time only"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload, "nodes_per_gpu": N, "directed_edges_per_gpu": E, "mp_layers": L,
                           "hidden": H, "parallelism": f"dp{world}",
                           "timing": "inputs (activations ~4 GB per step) larger than L2; no explicit flush",
                           "launch": "cuda-graph replay of the whole step" if graphed else "eager launches"},
                "train_steps_per_s": args.steps / (ms_res * 1e-3),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "pipeline": "next batch copied (pinned host -> device, side stream) during the current step; loss .item() every step",
                        "ms_per_step": ms_e2e / args.steps, "train_steps_per_s": args.steps / (ms_e2e * 1e-3)},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "loss_last": loss_host[-1]}
        if world == 1:
            # secondary figure (outside every timed region above): autoregressive roll-out on ONE mesh of the same
            # shape, per-frame step replayed from a CUDA graph -- the reference's validation loop (R1)
            try:
                one = cylinder_flow_batch(1, seed=0).to(dev)
                tr.rollout([one] * 5)
                torch.cuda.synchronize()
                t0 = time.time()
                tr.rollout([one] * 100)
                torch.cuda.synchronize()
                line["rollout"] = {"frames_per_s": 100 / (time.time() - t0), "nodes": int(one.x.shape[0]),
                                   "directed_edges": int(one.edge_index.shape[1]), "launch": "cuda-graph replay" if graphed else "eager"}
            except Exception as exc:       # never let the side figure break the contract line
                line["rollout"] = {"error": str(exc)[:200]}
        if world == 1 and not args.no_secondary:
            try:
                line["other_configs"] = secondary_configs(dev)
            except Exception as exc:
                line["other_configs"] = {"error": str(exc)[:300]}
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample (about 10-30 s of CPU work): 8 of the 32 graphs per step, 1 warm-up + 2 timed steps
            cb = run_reference(args, sample_graphs=8, budget_s=30.0, max_steps=2)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    part = None
    if world > 1 and not args.no_partition:
        # release the headline trainer (activations, captured graphs) first: the partition block's 1-GPU reference step
        # wants the whole device
        import gc
        from graphphysics_b200 import graph as _g
        tr = resident = None
        _g._PERSISTENT.clear()
        _g._CACHE.clear()
        gc.collect()
        torch.cuda.empty_cache()
        try:
            # the mesh grows with the job (each line carries its own 1-GPU reference on the same mesh): 373k nodes at 2
            # GPUs, 681k at 4, the 1M-node / 13.8M-edge mesh of BASELINE.json configs[4] at 8 -- when the unpartitioned
            # reference step (~145 GiB of activations at 1M nodes) fits next to what this process already holds
            side = args.partition_side or {2: 72, 4: 88}.get(world, 100 if world >= 8 else 72)
            free_gib = torch.cuda.mem_get_info(dev)[0] / 2 ** 30
            if side >= 100 and free_gib < 160:
                side = 88
            part = partition_block(dev, pg, world, rank, side=side)
        except Exception as exc:
            part = {"error": str(exc)[:300]}
    if rank == 0 and part is not None:
        line["partition"] = part
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # The captured steps hold NCCL work; tearing the communicator down under them can block, and
        # nothing is left to do: synchronise, then leave without the teardown.
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
