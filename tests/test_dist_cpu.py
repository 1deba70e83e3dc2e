"""CPU: multi-rank host logic with world_size 2 on the gloo backend -- gradient averaging, global
normaliser statistics, node partition + halo maps (bit-exact against the oracle) and the halo
exchange itself."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gp_oracle as O


def _run(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run, args=(world, port, fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _mesh():
    pos, tris = O.grid_tri_mesh(14, 9, jitter=0.3, seed=1)
    return pos, O.face_to_edge(tris, len(pos))


def test_partition_and_halo_maps_bit_exact():
    from graphphysics_b200.dist.partition import build_local_graphs, partition_nodes
    pos, ei = _mesh()
    for parts in (2, 4, 8):
        owner = partition_nodes(pos, parts)
        assert np.array_equal(owner, O.partition_nodes(pos, parts))
        counts = np.bincount(owner, minlength=parts)
        assert counts.sum() == len(pos) and counts.max() - counts.min() <= 1      # every node exactly once, balanced
        lgs, ref = build_local_graphs(ei, owner, parts), O.halo_maps(ei, owner, parts)
        kept = np.sort(np.concatenate([lg.edge_ids for lg in lgs]))
        assert np.array_equal(kept, np.arange(ei.shape[1]))                         # no edge dropped or duplicated
        for lg, r in zip(lgs, ref):
            assert np.array_equal(lg.owned, r["owned"]) and np.array_equal(lg.ghosts, r["ghosts"])
            assert np.array_equal(lg.edge_index_local, r["edge_index_local"])
            assert set(lg.send) == set(r["send"]) and set(lg.recv) == set(r["recv"])
            for q in lg.send:
                assert np.array_equal(lg.send[q], r["send"][q])
            for q in lg.recv:
                assert np.array_equal(lg.recv[q], r["recv"][q])
            assert (lg.edge_index_local[1] < lg.num_owned).all()                    # receivers are owned
        from graphphysics_b200.dist.partition import build_local_graph
        for p in range(parts):                                                       # the O(E) per-rank builder gives the same maps
            one = build_local_graph(ei, owner, parts, p)
            assert np.array_equal(one.owned, lgs[p].owned) and np.array_equal(one.ghosts, lgs[p].ghosts)
            assert np.array_equal(one.edge_ids, lgs[p].edge_ids) and np.array_equal(one.edge_index_local, lgs[p].edge_index_local)
            assert set(one.send) == set(lgs[p].send) and set(one.recv) == set(lgs[p].recv)
            assert all(np.array_equal(one.send[q], lgs[p].send[q]) for q in one.send)
            assert all(np.array_equal(one.recv[q], lgs[p].recv[q]) for q in one.recv)
        for p in range(parts):                                                       # send/recv lists mirror each other
            for q, idx in lgs[p].recv.items():
                glob_p = np.concatenate([lgs[p].owned, lgs[p].ghosts])[idx]
                glob_q = lgs[q].owned[lgs[q].send[p]]
                assert np.array_equal(glob_p, glob_q)


def _halo_worker(rank, world):
    from graphphysics_b200.dist.halo import HaloPlan
    from graphphysics_b200.dist.partition import build_local_graphs, partition_nodes
    pos, ei = _mesh()
    lg = build_local_graphs(ei, partition_nodes(pos, world), world)[rank]
    glob = np.concatenate([lg.owned, lg.ghosts])
    feat = torch.arange(len(pos), dtype=torch.float32)[:, None] * torch.tensor([[1.0, 10.0, 100.0]])
    x = feat[glob].clone()
    x[lg.num_owned:] = -1.0                                   # ghosts stale
    HaloPlan(lg, world, "cpu").exchange_(x)
    return bool(torch.equal(x, feat[glob]))


def test_halo_exchange_two_ranks():
    assert _spawn(_halo_worker) == [True, True]


def _halo_grad_worker(rank, world):
    """exchange_grad_ is the transpose of exchange_:  sum_ranks <exchange(x), g>  ==  sum_ranks <x_owned-part, exchange_grad(g)>
    for integer-valued (exactly representable) x and g, with the ghost rows of x treated as dead inputs."""
    import torch.distributed as dist
    from graphphysics_b200.dist.halo import HaloPlan
    from graphphysics_b200.dist.partition import build_local_graphs, partition_nodes
    pos, ei = _mesh()
    lg = build_local_graphs(ei, partition_nodes(pos, world), world)[rank]
    plan = HaloPlan(lg, world, "cpu")
    gen = torch.Generator().manual_seed(100 + rank)
    x = torch.randint(-8, 9, (lg.num_local, 3), generator=gen).float()
    g = torch.randint(-8, 9, (lg.num_local, 3), generator=gen).float()
    y = x.clone()
    plan.exchange_(y)                                   # y = A x  (ghost rows of x do not reach y)
    lhs = (y * g).sum()
    gt = g.clone()
    plan.exchange_grad_(gt)                             # gt = A^T g
    rhs = (x * gt).sum()
    both = torch.stack([lhs, rhs])
    dist.all_reduce(both)
    ghosts_zero = bool((gt[lg.num_owned:] == 0).all())
    return float(both[0]), float(both[1]), ghosts_zero


def test_halo_gradient_exchange_is_the_transpose():
    for lhs, rhs, ghosts_zero in _spawn(_halo_grad_worker):
        assert lhs == rhs and ghosts_zero


def _ddp_worker(rank, world):
    from graphphysics_b200.dist.ddp import accumulate_normalizers_globally, allreduce_mean_, broadcast_
    from graphphysics_b200.graph import Data
    from graphphysics_b200.models.simulator import Simulator
    g = torch.full((7,), float(rank + 1))
    allreduce_mean_(g)
    p = torch.full((3,), float(rank))
    broadcast_(p)
    torch.manual_seed(0)
    xs = [torch.cat([torch.randn(5, 2), torch.zeros(5, 2)], 1) for _ in range(world)]
    ys = [torch.randn(5, 2) for _ in range(world)]
    eas = [torch.randn(8, 3) for _ in range(world)]
    idx = dict(feature_index_start=0, feature_index_end=2, output_index_start=0, output_index_end=2, node_type_index=2)
    sim = Simulator(11, 3, 2, model=torch.nn.Identity(), device=torch.device("cpu"), **idx)
    accumulate_normalizers_globally(sim, Data(x=xs[rank], y=ys[rank], edge_attr=eas[rank]))
    ref = Simulator(11, 3, 2, model=torch.nn.Identity(), device=torch.device("cpu"), **idx)
    ref._build_input_graph(Data(x=torch.cat(xs), y=torch.cat(ys), edge_attr=torch.cat(eas),
                                edge_index=torch.zeros(2, 0, dtype=torch.long)), True)
    same = all(torch.allclose(a, b, rtol=1e-6, atol=1e-6) for a, b in zip(sim.state_dict().values(), ref.state_dict().values()))
    return (g.tolist(), p.tolist(), same)


def test_ddp_helpers_two_ranks():
    out = _spawn(_ddp_worker)
    for g, p, same in out:
        assert g == [1.5] * 7 and p == [0.0] * 3 and same
