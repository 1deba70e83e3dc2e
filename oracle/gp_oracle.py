"""CPU oracle for the graph-physics message-passing hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product (``graphphysics_b200``)
never does: it fails loudly when its CUDA library is missing.

This is a plain-PyTorch (CPU, fp32/fp64) restatement of the reference algorithm, each
function citing the ``/root/reference`` file:line it follows.  Parity status: PINNED --
``oracle/make_golden.py`` imports the unmodified reference modules (through
``oracle/ref_shim.py``) in the build container, runs them on seeded inputs, and stores
inputs / weights / outputs / loss / gradients under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those vectors (exact mode), and
against the integer goldens the reference's own tests hold (11 070 edges, 32 638 2-hop
edges, node counts).

Two arithmetic modes:
  * ``mode=None``   -- the reference's arithmetic (fp32 or fp64 everywhere).
  * ``mode="bf16"`` -- the *kernel specification*: same algorithm with values rounded to
    bf16 at exactly the points where the CUDA kernels round (MMA operands, tensors stored
    in HBM as bf16, gradient operands of the backward MMAs) and the first edge-MLP layer
    evaluated in its distributive form  W1e.e + (W1d.x)[dst] + (W1s.x)[src]  (SURVEY §7).
    Accumulation is fp32/fp64, so a kernel differs from it only by summation order.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

NODE_TYPE_SIZE = 9          # graphphysics/utils/nodetype.py:4-12
NORMAL, OBSTACLE, AIRFOIL, HANDLE, INFLOW, OUTFLOW, WALL_BOUNDARY = 0, 1, 2, 3, 4, 5, 6


# --------------------------------------------------------------------------- rounding helpers
class _RoundSTE(torch.autograd.Function):
    """Round the value to bf16 (RNE), pass the gradient through unchanged."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _GradRound(torch.autograd.Function):
    """Identity in forward; rounds the incoming gradient to bf16 (backward MMA operand)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def rnd(x: torch.Tensor, mode: Optional[str]) -> torch.Tensor:
    return _RoundSTE.apply(x) if mode == "bf16" else x


def grad_rnd(x: torch.Tensor, mode: Optional[str]) -> torch.Tensor:
    return _GradRound.apply(x) if mode == "bf16" else x


def linear(x, w, b, mode):
    """nn.Linear (y = x W^T + b) with bf16 MMA operands in kernel mode; the gradient that
    reaches the product is rounded too (it is the A operand of dgrad and wgrad)."""
    y = F.linear(rnd(x, mode), rnd(w, mode))
    if b is not None:
        y = y + b
    return grad_rnd(y, mode)        # delta (bf16) feeds dgrad, wgrad and the bias gradient alike


# --------------------------------------------------------------------------- layers
class _RmsNormKernel(torch.autograd.Function):
    """RMSNorm with the kernel's backward arithmetic: the scale gradient is the column sum of
    bf16(g * x/(rms+eps)) -- that product is staged as a bf16 MMA operand (delta^T . 1)."""

    @staticmethod
    def forward(ctx, x, scale, eps):
        d = x.shape[-1]
        rms = x.norm(2, dim=-1, keepdim=True) / math.sqrt(d)
        s = 1.0 / (rms + eps)
        ctx.save_for_backward(x, scale, s, rms)
        return scale * (x * s)

    @staticmethod
    def backward(ctx, g):
        x, scale, s, rms = ctx.saved_tensors
        d = x.shape[-1]
        q = (g * x * s).to(torch.bfloat16).to(g.dtype)
        dscale = q.reshape(-1, d).sum(0)
        dot = (g * scale * x).sum(-1, keepdim=True)
        coef = torch.where(rms > 0, dot * s * s / (rms * d), torch.zeros_like(dot))
        return scale * g * s - coef * x, dscale, None


def rms_norm(x: torch.Tensor, scale: torch.Tensor, eps: float = 1e-8, mode: Optional[str] = None) -> torch.Tensor:
    """RMSNorm.forward, full-vector branch (layers.py:104-129): scale * x / (||x||/sqrt(d) + eps)."""
    if mode == "bf16":
        return _RmsNormKernel.apply(x, scale, eps)
    d = x.shape[-1]
    rms = x.norm(2, dim=-1, keepdim=True) / math.sqrt(d)
    return scale * (x / (rms + eps))


_ACT = {"relu": F.relu, "gelu": F.gelu, "silu": F.silu}


def mlp(x, sd: Dict[str, torch.Tensor], prefix: str, nb_layers: int = 4, layer_norm: bool = True,
        act: str = "relu", mode: Optional[str] = None, first_pre: Optional[torch.Tensor] = None,
        preacts: Optional[list] = None):
    """build_mlp (layers.py:163-210): Linear,act,[Linear,act]*(nb-2),Linear,[RMSNorm].
    Sequential indices: linear i sits at 2*i, the norm at 2*nb-1.
    ``first_pre`` (kernel mode only) replaces the first layer's product with a precomputed
    pre-activation (bias excluded).  ``preacts``, if a list, receives the ReLU inputs (tests use
    it to keep inputs away from the kink, where summation order would pick the subgradient)."""
    h = x
    for i in range(nb_layers):
        w, b = sd[f"{prefix}.{2 * i}.weight"], sd[f"{prefix}.{2 * i}.bias"]
        if i == 0 and first_pre is not None:
            h = grad_rnd(first_pre + b, mode)
        else:
            h = linear(h, w, b, mode)
        if i < nb_layers - 1:
            if preacts is not None:
                preacts.append(h.detach())
            h = _ACT[act](h)
    if layer_norm:
        h = rms_norm(h, sd[f"{prefix}.{2 * nb_layers - 1}.scale"], mode=mode)
    return h


def graph_net_block(x, e, src, dst, sd, prefix: str, mode: Optional[str] = None):
    """GraphNetBlock.forward (layers.py:989-1102).
    edge input order [e, x[dst], x[src]] (layers.py:1016-1018, 1058); identity message summed
    at edge_index[1] (layers.py:926, 1031-1037; PyG propagate restated as index_add_);
    node input [x, agg] (layers.py:1100-1102); residuals (layers.py:1039-1040)."""
    N, H = x.shape
    if mode == "bf16":
        w1 = sd[f"{prefix}.edge_block.0.weight"]
        w1e, w1d, w1s = w1[:, :H], w1[:, H:2 * H], w1[:, 2 * H:]
        pd = rnd(linear(x, w1d, None, mode), mode)      # per-node projections, stored bf16
        ps = rnd(linear(x, w1s, None, mode), mode)
        # delta_1 (the gradient of this sum) is one bf16 tensor in the kernel: it feeds dE,
        # the receiver/sender segment sums and dW1e alike.
        pre = F.linear(rnd(e, mode), rnd(w1e, mode)) + pd[dst] + ps[src]
        e_upd = mlp(None, sd, f"{prefix}.edge_block", mode=mode, first_pre=pre)
        # kernel sums bf16(e_upd) in fp32; in the backward the gathered d agg rows are staged as bf16
        agg = torch.zeros_like(x).index_add_(0, dst, grad_rnd(rnd(e_upd, mode), mode))
        # node MLP, first layer split the same way: W1 = [W1x | W1a] over [x, agg]
        wn = sd[f"{prefix}.node_block.0.weight"]
        q = rnd(linear(x, wn[:, :H], None, mode), mode)
        pre_n = F.linear(rnd(agg, mode), rnd(wn[:, H:], mode)) + q
        x_upd = mlp(None, sd, f"{prefix}.node_block", mode=mode, first_pre=pre_n)
    else:
        e_upd = mlp(torch.cat([e, x[dst], x[src]], dim=-1), sd, f"{prefix}.edge_block")
        agg = torch.zeros_like(x).index_add_(0, dst, e_upd)
        x_upd = mlp(torch.cat([x, agg], dim=-1), sd, f"{prefix}.node_block")
    # the kernels round the normalised update to bf16 once; that value feeds agg and the residual
    return rnd(x + rnd(x_upd, mode), mode), rnd(e + rnd(e_upd, mode), mode)


def epd_forward(sd, x_in, edge_attr, edge_index, num_layers: int, mode: Optional[str] = None,
                prefix: str = "", only_processor: bool = False):
    """EncodeProcessDecode.forward (processors.py:162-215)."""
    src, dst = edge_index[0], edge_index[1]
    if only_processor:
        x, e = x_in, edge_attr
    else:
        x = rnd(mlp(x_in, sd, f"{prefix}nodes_encoder", mode=mode), mode)
        e = rnd(mlp(edge_attr, sd, f"{prefix}edges_encoder", mode=mode), mode)
    for i in range(num_layers):
        x, e = graph_net_block(x, e, src, dst, sd, f"{prefix}processor_list.{i}", mode)
    if only_processor:
        return x
    return mlp(x, sd, f"{prefix}decode_module", layer_norm=False, mode=mode)


def sparse_attention(q, k, v, row, col, num_nodes: int):
    """scaled_dot_product_attention on the DGL branch (layers.py:493-561).
    q,k,v: (N, d, Hh) with the head axis innermost (layers.py:673-675).  For every stored
    (i=row, j=col): s = sum_d q[i,d,h] k[j,d,h] / sqrt(d) (bsddmm, layers.py:509-516); softmax
    over the entries of row i, per head (layers.py:517); y[i] = sum_j a_ij v[j] (bspmm, 554)."""
    d = q.shape[1]
    s = (q[row] / math.sqrt(d) * k[col]).sum(dim=1)                    # (E, Hh)
    smax = torch.full((num_nodes, s.shape[1]), -float("inf"), dtype=s.dtype)
    smax = smax.scatter_reduce(0, row[:, None].expand_as(s), s, reduce="amax", include_self=True)
    p = torch.exp(s - smax[row])
    den = torch.zeros((num_nodes, s.shape[1]), dtype=s.dtype).index_add_(0, row, p)
    a = p / den[row]
    y = torch.zeros_like(q).index_add_(0, row, a[:, None, :] * v[col])
    return y


def attention_values(q, k, row, col, num_nodes: int):
    """scaled_query_key_softmax on the DGL branch (layers.py:493-522): the values of the sparse softmax, one row per stored
    entry (in the order of row / col) and one column per head -- what return_attention=True exposes as attn.val."""
    d = q.shape[1]
    s = (q[row] / math.sqrt(d) * k[col]).sum(dim=1)
    smax = torch.full((num_nodes, s.shape[1]), -float("inf"), dtype=s.dtype)
    smax = smax.scatter_reduce(0, row[:, None].expand_as(s), s, reduce="amax", include_self=True)
    p = torch.exp(s - smax[row])
    den = torch.zeros((num_nodes, s.shape[1]), dtype=s.dtype).index_add_(0, row, p)
    return p / den[row]


def attention(x, row, col, sd, prefix: str, num_heads: int, mode: Optional[str] = None):
    """Attention.forward (layers.py:637-697), no RoPE / gate."""
    N, H = x.shape
    d = H // num_heads
    q = linear(x, sd[f"{prefix}.q_proj.weight"], sd.get(f"{prefix}.q_proj.bias"), mode)
    k = linear(x, sd[f"{prefix}.k_proj.weight"], sd.get(f"{prefix}.k_proj.bias"), mode)
    v = linear(x, sd[f"{prefix}.v_proj.weight"], sd.get(f"{prefix}.v_proj.bias"), mode)
    q, k, v = (rnd(t, mode).reshape(N, d, num_heads) for t in (q, k, v))
    y = sparse_attention(q, k, v, row, col, N).reshape(N, H)
    return linear(rnd(y, mode), sd[f"{prefix}.proj.weight"], sd.get(f"{prefix}.proj.bias"), mode)


def transformer_block(x, row, col, sd, prefix: str, num_heads: int, mode: Optional[str] = None):
    """Transformer.forward (layers.py:766-819) with build_gated_mlp (layers.py:252-278, 213-249):
    x += Attn(norm1(x));  x += W3( GELU(W1 n) * (W2 n) ),  n = RMSNorm_g(RMSNorm_2(x)).
    Kernel mode: the residual stream x stays fp32 (both adds ride in a GEMM epilogue); the norm outputs, q / k / v / y and
    the gate output are bf16 (MMA operands); the scale gradients are plain fp32 column sums (gp_rmsnorm_bwd)."""
    n1 = rms_norm(x, sd[f"{prefix}.norm1.scale"])
    x = x + attention(rnd(n1, mode), row, col, sd, f"{prefix}.attention", num_heads, mode)
    n2 = rms_norm(rms_norm(x, sd[f"{prefix}.norm2.scale"]), sd[f"{prefix}.gated_mlp.0.scale"])
    n2 = rnd(n2, mode)
    left = F.gelu(linear(n2, sd[f"{prefix}.gated_mlp.1.linear1.weight"], sd[f"{prefix}.gated_mlp.1.linear1.bias"], mode))
    right = linear(n2, sd[f"{prefix}.gated_mlp.1.linear2.weight"], sd[f"{prefix}.gated_mlp.1.linear2.bias"], mode)
    g = rnd(left * right, mode)
    out = linear(g, sd[f"{prefix}.gated_mlp.2.weight"], sd[f"{prefix}.gated_mlp.2.bias"], mode)
    return x + out


def dense_mlp(x, sd, prefix: str, layer_norm: bool = True, mode: Optional[str] = None):
    """build_mlp as the Transformer path runs it (graphphysics_b200/dense.py mlp4): one GEMM per Linear, hidden
    activations stored as bf16, fp32 output and a plain fp32 RMSNorm."""
    h = x
    for i in range(4):
        h = linear(h, sd[f"{prefix}.{2 * i}.weight"], sd[f"{prefix}.{2 * i}.bias"], mode)
        if i < 3:
            h = rnd(F.relu(h), mode)
    return rms_norm(h, sd[f"{prefix}.7.scale"]) if layer_norm else h


def etd_forward(sd, x_in, edge_index, num_layers: int, num_heads: int, mode: Optional[str] = None, prefix: str = ""):
    """EncodeTransformDecode.forward, DGL branch (processors.py:338-384): adjacency rows are
    edge_index[0], columns edge_index[1] (processors.py:366), no self loops added."""
    row, col = edge_index[0], edge_index[1]
    x = dense_mlp(x_in, sd, f"{prefix}nodes_encoder", mode=mode)
    for i in range(num_layers):
        x = transformer_block(x, row, col, sd, f"{prefix}processor_list.{i}", num_heads, mode)
    return dense_mlp(x, sd, f"{prefix}decode_module", layer_norm=False, mode=mode)


class Normalizer:
    """Normalizer (layers.py:281-408): running sum / sum of squares / count, frozen after
    max_accumulations calls; (x - mean) / max(std, eps)."""

    def __init__(self, size: int, max_accumulations: int = 10 ** 5, std_epsilon: float = 1e-8, dtype=torch.float32):
        self.max_acc = max_accumulations
        self.eps = std_epsilon
        self.acc_count = torch.zeros((), dtype=dtype)
        self.num_acc = torch.zeros((), dtype=dtype)
        self.acc_sum = torch.zeros((1, size), dtype=dtype)
        self.acc_sum_sq = torch.zeros((1, size), dtype=dtype)

    def load(self, sd, prefix):
        self.acc_count = sd[f"{prefix}._acc_count"].clone()
        self.num_acc = sd[f"{prefix}._num_accumulations"].clone()
        self.acc_sum = sd[f"{prefix}._acc_sum"].clone()
        self.acc_sum_sq = sd[f"{prefix}._acc_sum_squared"].clone()

    def mean(self):
        return self.acc_sum / torch.clamp(self.acc_count, min=1.0)

    def std(self):
        var = self.acc_sum_sq / torch.clamp(self.acc_count, min=1.0) - self.mean() ** 2
        return torch.clamp(torch.sqrt(torch.clamp(var, min=0.0)), min=self.eps)

    def __call__(self, x, accumulate: bool = True):
        if accumulate and float(self.num_acc) < self.max_acc:       # layers.py:345-349
            d = x.detach()
            self.acc_sum = self.acc_sum + d.sum(0, keepdim=True)
            self.acc_sum_sq = self.acc_sum_sq + (d ** 2).sum(0, keepdim=True)
            self.acc_count = self.acc_count + d.shape[0]
            self.num_acc = self.num_acc + 1
        return (x - self.mean()) / self.std()

    def inverse(self, y):
        return y * self.std() + self.mean()


def build_node_features(x_raw, feat_s: int, feat_e: int, node_type_index: int):
    """Simulator._get_one_hot_type / _build_node_features (simulator.py:112-143)."""
    one_hot = F.one_hot(x_raw[:, node_type_index].long(), NODE_TYPE_SIZE).to(x_raw.dtype)
    return torch.cat([x_raw[:, feat_s:feat_e], one_hot], dim=1)


def simulator_forward(model_fn, norms: Dict[str, Optional[Normalizer]], x_raw, y, edge_attr, index: Dict[str, int],
                      training: bool):
    """Simulator.forward (simulator.py:145-217).  ``model_fn(node_feat, edge_feat) -> net_out``.
    Returns (network_output, target_delta_normalized, outputs-or-None)."""
    pre_target = x_raw[:, index["output_index_start"]:index["output_index_end"]]
    target_norm = norms["output"](y - pre_target, training)
    nf = build_node_features(x_raw, index["feature_index_start"], index["feature_index_end"], index["node_type_index"])
    nf = norms["node"](nf, training)
    ef = norms["edge"](edge_attr, training) if norms.get("edge") is not None else edge_attr
    net_out = model_fn(nf, ef)
    if training:
        return net_out, target_norm, None
    return net_out, target_norm, pre_target + norms["output"].inverse(net_out)


def l2_loss(target, network_output, node_type, masks: Sequence[int] = (NORMAL, OUTFLOW)):
    """L2Loss.forward + _prepare_mask_for_loss (loss.py:19-75): mean over masked rows x out dims."""
    mask = torch.zeros_like(node_type, dtype=torch.bool)
    for m in masks:
        mask |= node_type == m
    return ((network_output - target) ** 2)[mask].mean()


def cosine_warmup_factor(step_index: int, warmup: int, max_iters: int, min_lr_factor: float = 1e-3) -> float:
    """CosineWarmupScheduler.get_lr_factor (scheduler.py:51-67); step_index is ``last_epoch``."""
    epoch = step_index + 1
    f = 0.5 * (1 + np.cos(np.pi * epoch / max_iters))
    if epoch <= warmup:
        f *= epoch * 1.0 / warmup
    return float(max(f, min_lr_factor))


def boundary_mask(node_type):
    """build_mask (lightning_module.py:27-35): True where the node is NOT NORMAL/OUTFLOW."""
    return ~((node_type == NORMAL) | (node_type == OUTFLOW))


def rollout(step_fn, frames_x: List[torch.Tensor], frames_y: List[torch.Tensor], out_s: int, out_e: int,
            node_type_index: int):
    """_make_prediction / validation_step (lightning_module.py:375-456): autoregressive roll-out
    with ground-truth overwrite on boundary nodes; returns predictions, 1-step RMSE, roll-out RMSE."""
    last = None
    preds = []
    for x_raw, y in zip(frames_x, frames_y):
        x_raw = x_raw.clone()
        if last is not None:
            x_raw[:, out_s:out_e] = last
        out = step_fn(x_raw, y).clone()
        m = boundary_mask(x_raw[:, node_type_index])
        out[m] = y[m]
        last = out
        preds.append(out)
    p, t = torch.cat(preds), torch.cat(frames_y)
    rmse_1 = torch.sqrt(((preds[0] - frames_y[0]) ** 2).mean()).item()
    rmse_all = torch.sqrt(((p - t) ** 2).mean()).item()
    return preds, rmse_1, rmse_all


# --------------------------------------------------------------------------- graph construction
def tetra_to_faces(tetra: np.ndarray) -> np.ndarray:
    """torch_graph.py:194-210: every tetrahedron contributes its 4 triangles. tetra: (T,4) -> (4T,3)."""
    t = np.asarray(tetra)
    return np.concatenate([t[:, [0, 1, 2]], t[:, [0, 1, 3]], t[:, [0, 2, 3]], t[:, [1, 2, 3]]], axis=0)


def face_to_edge(faces: np.ndarray, num_nodes: int) -> np.ndarray:
    """PyG T.FaceToEdge (torch-geometric 2.6.1, used at preprocessing.py:410-424):
    edges (f0,f1),(f1,f2),(f2,f0) of every triangle, made undirected and coalesced, i.e.
    unique directed pairs sorted by (row, col).  faces: (F,3).  Returns int64 (2,E)."""
    f = np.asarray(faces, dtype=np.int64).T                       # (3, F) like data.face
    ei = np.concatenate([f[:2], f[1:], f[::2]], axis=1)
    both = np.concatenate([ei, ei[::-1]], axis=1)
    key = np.unique(both[0] * num_nodes + both[1])
    return np.stack([key // num_nodes, key % num_nodes])


def edge_features(pos: np.ndarray, edge_index: np.ndarray) -> np.ndarray:
    """T.Cartesian(norm=False) then T.Distance(norm=False) (preprocessing.py:16-23):
    [pos[row]-pos[col], ||pos[col]-pos[row]||]  ->  (E, dim+1)."""
    row, col = edge_index
    cart = pos[row] - pos[col]
    dist = np.linalg.norm(pos[col] - pos[row], axis=-1, keepdims=True)
    return np.concatenate([cart, dist], axis=-1).astype(pos.dtype)


def khop_edges(edge_index: np.ndarray, num_nodes: int, k: int) -> np.ndarray:
    """torch_graph.py:14-54 (k-hop via powers of the adjacency, self loops removed)."""
    import scipy.sparse as sp
    a = sp.coo_matrix((np.ones(edge_index.shape[1]), (edge_index[0], edge_index[1])), shape=(num_nodes, num_nodes)).tocsr()
    a.data[:] = 1
    acc, p = a.copy(), a.copy()
    for _ in range(k - 1):
        p = (p @ a)
        p.data[:] = 1
        acc = acc + p
    acc.setdiag(0)
    acc.eliminate_zeros()
    c = acc.tocoo()
    key = np.unique(c.row.astype(np.int64) * num_nodes + c.col)
    return np.stack([key // num_nodes, key % num_nodes])


def csr_by_receiver(edge_index: np.ndarray, num_nodes: int):
    """Kernel-side graph layout: stable sort of the edges by receiver (edge_index[1]) and by
    sender.  Returns dict(perm_dst, rowptr_dst, perm_src, rowptr_src) (int32; perm maps sorted
    position -> original edge id)."""
    src, dst = edge_index[0], edge_index[1]
    perm_dst = np.argsort(dst, kind="stable").astype(np.int32)
    perm_src = np.argsort(src, kind="stable").astype(np.int32)
    rp_dst = np.zeros(num_nodes + 1, np.int32)
    rp_src = np.zeros(num_nodes + 1, np.int32)
    np.cumsum(np.bincount(dst, minlength=num_nodes), out=rp_dst[1:])
    np.cumsum(np.bincount(src, minlength=num_nodes), out=rp_src[1:])
    return dict(perm_dst=perm_dst, rowptr_dst=rp_dst, perm_src=perm_src, rowptr_src=rp_src)


# --------------------------------------------------------------------------- node partition + halo maps
def partition_nodes(pos: np.ndarray, num_parts: int) -> np.ndarray:
    """Deterministic k-way node partition by recursive coordinate bisection (METIS is not
    available).  Splits the widest axis at the balanced rank, ties broken by node
    id.  num_parts must be a power of two.  Returns owner[n] in [0, num_parts)."""
    n = pos.shape[0]
    owner = np.zeros(n, np.int32)

    def rec(ids, lo, parts):
        if parts == 1:
            owner[ids] = lo
            return
        p = pos[ids]
        axis = int(np.argmax(p.max(0) - p.min(0)))
        order = np.lexsort((ids, p[:, axis]))
        half = (len(ids) * (parts // 2)) // parts
        rec(ids[order[:half]], lo, parts // 2)
        rec(ids[order[half:]], lo + parts // 2, parts - parts // 2)

    rec(np.arange(n, dtype=np.int64), 0, num_parts)
    return owner


def halo_maps(edge_index: np.ndarray, owner: np.ndarray, num_parts: int):
    """Per-rank local graphs for node-partitioned message passing.
    An edge lives on the rank that owns its receiver.  Rank p's local node order is
    [owned nodes ascending by global id | ghost senders ascending by global id].
    Returns a list (per rank) of dicts:
      owned (global ids), ghosts (global ids), edge_ids (global edge ids kept, original order),
      edge_index_local (2,E_p) in local numbering,
      send[q] = local indices (into owned) of rows rank p must send to q,
      recv[q] = local indices (ghost slots, offset by len(owned)) filled from rank q.
    send lists on p and recv lists on q enumerate the same global ids in ascending order."""
    src, dst = edge_index[0], edge_index[1]
    out = []
    for p in range(num_parts):
        owned = np.nonzero(owner == p)[0].astype(np.int64)
        eids = np.nonzero(owner[dst] == p)[0].astype(np.int64)
        s = src[eids]
        ghosts = np.unique(s[owner[s] != p]).astype(np.int64)
        local_of = {}
        glob = np.concatenate([owned, ghosts])
        lut = np.full(owner.shape[0], -1, np.int64)
        lut[glob] = np.arange(len(glob))
        ei_local = np.stack([lut[s], lut[dst[eids]]])
        recv = {}
        for q in range(num_parts):
            if q == p:
                continue
            g = ghosts[owner[ghosts] == q]
            if len(g):
                recv[q] = lut[g].astype(np.int32)
        out.append(dict(owned=owned, ghosts=ghosts, edge_ids=eids, edge_index_local=ei_local.astype(np.int64),
                        recv=recv, send={}, lut=lut))
    for p in range(num_parts):
        for q, idx in out[p]["recv"].items():
            g = np.concatenate([out[p]["owned"], out[p]["ghosts"]])[idx]
            out[q]["send"][p] = out[q]["lut"][g].astype(np.int32)
    for d in out:
        d.pop("lut")
    return out


# --------------------------------------------------------------------------- synthetic meshes
def grid_tri_mesh(nx: int, ny: int, jitter: float = 0.0, seed: int = 0, hole: Optional[Tuple[float, float, float]] = None):
    """Structured triangulated rectangle (optionally jittered, with a circular hole):
    the CylinderFlow-shaped synthetic mesh of SURVEY §8d.  Returns pos (N,2) float32, tris (T,3)."""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.linspace(0, 1.6, nx), np.linspace(0, 0.41, ny), indexing="ij")
    pos = np.stack([xs.ravel(), ys.ravel()], -1)
    idx = np.arange(nx * ny).reshape(nx, ny)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel()
    tris = np.concatenate([np.stack([a, b, c], -1), np.stack([b, d, c], -1)])
    if jitter:
        inner = np.ones(nx * ny, bool).reshape(nx, ny)
        inner[0], inner[-1], inner[:, 0], inner[:, -1] = False, False, False, False
        pos = pos + inner.ravel()[:, None] * rng.uniform(-jitter, jitter, pos.shape) * np.array([1.6 / nx, 0.41 / ny])
    if hole is not None:
        cx, cy, r = hole
        keep = ((pos[:, 0] - cx) ** 2 + (pos[:, 1] - cy) ** 2) > r * r
        tris = tris[keep[tris].all(1)]
        used = np.zeros(len(pos), bool)
        used[tris.ravel()] = True
        remap = np.cumsum(used) - 1
        pos, tris = pos[used], remap[tris]
    return pos.astype(np.float32), tris.astype(np.int64)


def grid_tet_mesh(nx: int, ny: int, nz: int):
    """Structured box split into 6 tetrahedra per cell. Returns pos (N,3) float32, tets (T,4)."""
    xs, ys, zs = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny), np.linspace(0, 1, nz), indexing="ij")
    pos = np.stack([xs.ravel(), ys.ravel(), zs.ravel()], -1)
    idx = np.arange(nx * ny * nz).reshape(nx, ny, nz)
    c = [idx[i:nx - 1 + i, j:ny - 1 + j, k:nz - 1 + k].ravel() for i in (0, 1) for j in (0, 1) for k in (0, 1)]
    v000, v001, v010, v011, v100, v101, v110, v111 = c
    tets = np.concatenate([
        np.stack([v000, v100, v110, v111], -1), np.stack([v000, v110, v010, v111], -1),
        np.stack([v000, v010, v011, v111], -1), np.stack([v000, v011, v001, v111], -1),
        np.stack([v000, v001, v101, v111], -1), np.stack([v000, v101, v100, v111], -1)])
    return pos.astype(np.float32), tets.astype(np.int64)


# --------------------------------------------------------------------------- preprocessing around the path
def world_edges(edge_index: np.ndarray, world_pos: np.ndarray, node_type: np.ndarray, num_nodes: int, radius: float = 0.03) -> np.ndarray:
    """add_world_edges (preprocessing.py:92-140): cKDTree.query_pairs(radius) on the fp32 world positions (scipy, the
    reference's own third-party call: pairs i < j with distance <= radius), kept when one end is OBSTACLE and the other
    NORMAL, concatenated in front of the mesh edges and passed through to_undirected (both directions, coalesced,
    sorted by (row, col)).  Returns int64 (2, E)."""
    from scipy.spatial import cKDTree
    pairs = cKDTree(np.asarray(world_pos)).query_pairs(radius, output_type="ndarray").T.astype(np.int64)
    t0, t1 = node_type[pairs[0]], node_type[pairs[1]]
    keep = ((t0 == OBSTACLE) & (t1 == NORMAL)) | ((t0 == NORMAL) & (t1 == OBSTACLE))
    ei = np.concatenate([pairs[:, keep], np.asarray(edge_index, dtype=np.int64)], axis=1)
    both = np.concatenate([ei, ei[::-1]], axis=1)
    key = np.unique(both[0] * np.int64(num_nodes) + both[1])
    return np.stack([key // num_nodes, key % num_nodes])


def world_pos_features(edge_attr: np.ndarray, world_pos: np.ndarray, edge_index: np.ndarray) -> np.ndarray:
    """add_world_pos_features (preprocessing.py:143-175): [edge_attr, wp[senders]-wp[receivers], ||.||_2], fp32."""
    s, r = edge_index
    rel = (world_pos[s] - world_pos[r]).astype(np.float32)
    back = (world_pos[r] - world_pos[s]).astype(np.float32)          # same magnitude; spelled like edge_features
    nrm = np.linalg.norm(back, axis=-1, keepdims=True).astype(np.float32)
    return np.concatenate([edge_attr.astype(np.float32), rel, nrm], -1)


def add_noise(x: np.ndarray, noise: np.ndarray, start: int, end: int, scale: float, node_type_index: int, t: Optional[float] = None):
    """add_noise (preprocessing.py:177-238) with the Gaussian draw passed in: x[:, start:end] += noise * scale on NORMAL
    nodes (scale(t) = 10 * scale * (1 + cos(pi t)) under the curriculum); fp32."""
    s = np.float32(10 * scale * (1 + math.cos(t * math.pi)) if t is not None else scale)
    out = x.copy()
    keep = (x[:, node_type_index] == NORMAL)[:, None]
    out[:, start:end] = x[:, start:end] + np.where(keep, noise.astype(np.float32) * s, np.float32(0))
    return out


# --------------------------------------------------------------------------- variant flags
def gated_mlp_seq(x, sd, prefix: str, act: str = "gelu", mode: Optional[str] = None):
    """build_gated_mlp (layers.py:252-278): RMSNorm -> GatedMLP (act(W1 n) * (W2 n), layers.py:213-249) -> Linear."""
    n = rnd(rms_norm(x, sd[f"{prefix}.0.scale"]), mode)
    left = _ACT[act](linear(n, sd[f"{prefix}.1.linear1.weight"], sd[f"{prefix}.1.linear1.bias"], mode))
    right = linear(n, sd[f"{prefix}.1.linear2.weight"], sd[f"{prefix}.1.linear2.bias"], mode)
    return linear(rnd(left * right, mode), sd[f"{prefix}.2.weight"], sd[f"{prefix}.2.bias"], mode)


def rope_rel(x_src, delta_pos, axes: int, base: float = 10000.0):
    """GraphNetBlock._apply_rope_rel (layers.py:1104-1149): per axis, `pair_count` (even, odd) channel pairs rotated by
    theta = delta[axis] * base^(-i / pair_count); the remaining channels pass through."""
    E, H = x_src.shape
    pc = H // (2 * axes)
    if pc == 0:
        return x_src
    inv = torch.pow(torch.tensor(base, dtype=torch.float32), -torch.arange(pc, dtype=torch.float32) / max(float(pc), 1.0))
    parts, start = [], 0
    for a in range(axes):
        seg = x_src[:, start:start + 2 * pc].reshape(E, pc, 2)
        theta = delta_pos[:, a].to(inv.dtype).unsqueeze(1) * inv.unsqueeze(0)
        c, s = torch.cos(theta).to(x_src.dtype), torch.sin(theta).to(x_src.dtype)
        ev, od = seg[..., 0], seg[..., 1]
        parts.append(torch.stack([ev * c - od * s, ev * s + od * c], -1).reshape(E, 2 * pc))
        start += 2 * pc
    return torch.cat(parts + [x_src[:, axes * 2 * pc:]], -1)


def graph_net_block_variant(x, e, src, dst, sd, prefix: str, *, act: str = "relu", gated_mlp: bool = False, gate: bool = False,
                            rope_axes: int = 0, rope_base: float = 10000.0, pos=None, phi=None, mode: Optional[str] = None,
                            nb_layers: int = 4, layer_norm: bool = True):
    """GraphNetBlock.forward with its constructor flags (layers.py:989-1102).  Kernel mode (`mode="bf16"`): GEMM operands
    and the gradients entering them rounded to bf16, everything else in the working precision -- the arithmetic of
    graphphysics_b200/variants.py."""
    x_i, x_j = x[dst], x[src]
    if rope_axes:
        x_j = rope_rel(x_j, pos[src, :rope_axes] - pos[dst, :rope_axes], rope_axes, rope_base)
    cat = torch.cat([e, x_i, x_j], -1)
    if gated_mlp:
        e_upd = gated_mlp_seq(cat, sd, f"{prefix}.edge_block", "silu" if act == "silu" else "gelu", mode)
    else:
        e_upd = dense_mlp_act(cat, sd, f"{prefix}.edge_block", act, mode, layer_norm, nb_layers)
    agg = torch.zeros_like(x).index_add_(0, dst, e_upd)
    if gate:
        logits = linear(x, sd[f"{prefix}.gate_proj.weight"], sd[f"{prefix}.gate_proj.bias"], mode)
        if phi is not None:
            logits = logits + phi.reshape(-1, 1).to(logits.dtype) * sd[f"{prefix}.gate_pos"].reshape(1, -1)
        agg = agg * torch.sigmoid(logits)
    cat_n = torch.cat([x, agg], -1)
    if gated_mlp:
        x_upd = gated_mlp_seq(cat_n, sd, f"{prefix}.node_block", "silu" if act == "silu" else "gelu", mode)
    else:
        x_upd = dense_mlp_act(cat_n, sd, f"{prefix}.node_block", act, mode, layer_norm, nb_layers)
    return x + x_upd, e + e_upd


def dense_mlp_act(x, sd, prefix: str, act: str = "relu", mode: Optional[str] = None, layer_norm: bool = True, nb_layers: int = 4):
    """build_mlp (layers.py:163-210) with a selectable activation and depth, one GEMM per Linear (hidden activations are
    MMA operands): Linear, act, ..., Linear [, RMSNorm] -- the norm is module 2 * nb_layers - 1 of the Sequential."""
    h = x
    for i in range(nb_layers):
        h = linear(h, sd[f"{prefix}.{2 * i}.weight"], sd[f"{prefix}.{2 * i}.bias"], mode)
        if i < nb_layers - 1:
            h = _ACT[act](h)
    return rms_norm(h, sd[f"{prefix}.{2 * nb_layers - 1}.scale"]) if layer_norm else h


def temporal_attention(h_prev, h_pred, row, col, sd, prefix: str, num_heads: int = 4, use_gate: bool = True, mode: Optional[str] = None):
    """TemporalAttention.forward (layers.py:861-887): q, v = Linear(h_pred), k = Linear(h_prev), heads innermost;
    scaled_dot_product_attention over the adjacency; out_proj; sigmoid gate MLP on [h_pred | h_prev]; residual onto
    h_prev; SiLU mixer on [h_corr | h_prev].  Kernel mode rounds GEMM operands only (q / k / v stay fp32)."""
    N, H = h_prev.shape
    d = H // num_heads
    lin = lambda x, n: linear(x, sd[f"{prefix}.{n}.weight"], sd[f"{prefix}.{n}.bias"], mode)
    q, k, v = lin(h_pred, "q_proj"), lin(h_prev, "k_proj"), lin(h_pred, "v_proj")
    y = sparse_attention(*(t.reshape(N, d, num_heads) for t in (q, k, v)), row, col, N).reshape(N, H)
    out = lin(y, "out_proj")
    if use_gate:
        out = torch.sigmoid(lin(F.silu(lin(torch.cat([h_pred, h_prev], -1), "gate.0")), "gate.2")) * out
    h_corr = h_prev + out
    return h_corr + lin(F.silu(lin(torch.cat([h_corr, h_prev], -1), "mixer.0")), "mixer.2")


def epd_forward_variant(sd, x_in, edge_attr, edge_index, num_layers: int, *, act: str = "relu", gated_mlp: bool = False,
                        gate: bool = False, rope_axes: int = 0, rope_base: float = 10000.0, pos=None, phi=None,
                        temporal: bool = False, mode: Optional[str] = None):
    """EncodeProcessDecode.forward with the variant flags (processors.py:162-215)."""
    src, dst = edge_index[0], edge_index[1]
    x = dense_mlp_act(x_in, sd, "nodes_encoder", act, mode)
    e = dense_mlp_act(edge_attr, sd, "edges_encoder", act, mode)
    prev_x = x
    for i in range(num_layers):
        prev_x = x
        x, e = graph_net_block_variant(x, e, src, dst, sd, f"processor_list.{i}", act=act, gated_mlp=gated_mlp, gate=gate,
                                       rope_axes=rope_axes, rope_base=rope_base, pos=pos, phi=phi if gate else None, mode=mode)
    if temporal:                                                          # processors.py:204-209
        x = temporal_attention(prev_x, x, src, dst, sd, "temporal_block", 4, True, mode)
    return dense_mlp_act(x, sd, "decode_module", act, mode, layer_norm=False)


def rope_nodes(q, k, pos, inv_freq):
    """_apply_rope_with_inv (layers.py:420-491): q, k (N, D, Hh); per position axis a and frequency i the channel pair
    (d_even, d_odd) = (a*2m + 2i, a*2m + 2i + 1) of every head is rotated by pos[:, a] * inv_freq[i]."""
    N, D, Hh = q.shape
    pd = pos.shape[1]
    m = D // (pd * 2)
    if m == 0:
        return q, k
    ang = pos[:, :pd].to(torch.float32).unsqueeze(-1) * inv_freq.to(torch.float32).view(1, 1, m)
    s, c = torch.sin(ang).to(q.dtype), torch.cos(ang).to(q.dtype)

    def ap(t):
        part = t[:, :pd * 2 * m, :].reshape(N, pd, m, 2, Hh)
        ev, od = part[..., 0, :], part[..., 1, :]
        rot = torch.stack((ev * c.unsqueeze(-1) - od * s.unsqueeze(-1), ev * s.unsqueeze(-1) + od * c.unsqueeze(-1)), 3).reshape(N, pd * 2 * m, Hh)
        return torch.cat([rot, t[:, pd * 2 * m:, :]], 1)

    return ap(q), ap(k)


def etd_forward_variant(sd, x_in, edge_index, num_layers: int, num_heads: int, *, act: str = "relu", gated_attention: bool = False,
                        rope: bool = False, rope_base: float = 10000.0, pos=None, temporal: bool = False, mode: Optional[str] = None):
    """EncodeTransformDecode.forward with use_gated_attention / use_rope_embeddings / SiLU (processors.py:338-384,
    layers.py:637-697, 766-819)."""
    row, col = edge_index[0], edge_index[1]
    x = dense_mlp_act(x_in, sd, "nodes_encoder", act, mode)
    N, H = x.shape
    d = H // num_heads
    prev_x = x
    for i in range(num_layers):
        prev_x = x
        p = f"processor_list.{i}"
        n1 = rnd(rms_norm(x, sd[f"{p}.norm1.scale"]), mode)
        q, k, v = (linear(n1, sd[f"{p}.attention.{w}_proj.weight"], sd.get(f"{p}.attention.{w}_proj.bias"), mode) for w in "qkv")
        q, k, v = (t.reshape(N, d, num_heads) for t in (q, k, v))
        if rope:
            m = d // max(pos.shape[1] * 2, 1)
            inv = torch.exp(-torch.arange(m, dtype=torch.float32) * (math.log(rope_base) / max(m, 1)))      # _make_inv_freq
            q, k = rope_nodes(q, k, pos, inv)
        q, k, v = (rnd(t, mode) for t in (q, k, v))
        y = sparse_attention(q, k, v, row, col, N)
        if gated_attention:
            gl = linear(n1, sd[f"{p}.attention.gate_proj.weight"], sd.get(f"{p}.attention.gate_proj.bias"), mode)
            y = y * torch.sigmoid(gl).reshape(N, d, num_heads)
        x = x + linear(rnd(y.reshape(N, H), mode), sd[f"{p}.attention.proj.weight"], sd.get(f"{p}.attention.proj.bias"), mode)
        x = x + gated_mlp_seq(rms_norm(x, sd[f"{p}.norm2.scale"]), sd, f"{p}.gated_mlp", "silu" if act == "silu" else "gelu", mode)
    if temporal:                                                          # processors.py:376-377
        x = temporal_attention(prev_x, x, row, col, sd, "temporal_block", num_heads, True, mode)
    return dense_mlp_act(x, sd, "decode_module", act, mode, layer_norm=False)
