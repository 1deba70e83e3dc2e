"""Execution engine of the encode-process-decode model on the CUDA kernels.

Owns what the kernels need and PyTorch modules do not have: one flat fp32 parameter buffer (the
nn.Parameters of the model become views into it, so `state_dict` keys and shapes stay those of
the reference), one flat gradient buffer written directly by the backward
kernels, the packed bf16 operand copies of every weight matrix, and the per-step workspaces.

Forward of one GraphNetBlock (graphphysics/models/layers.py:989-1102) is three launches:
    P            = x . [W1d ; W1s ; W1x]^T                (per-node halves of both first layers)
    e', agg      = edge kernel(e, P[dst], P[src])          (MLP + RMSNorm + residual + segment sum)
    x'           = node kernel(agg, P[self], x)            (MLP + RMSNorm + residual)
and the backward re-computes them stage by stage from the layer inputs (x, e), the projection P
and the one activation per MLP the forward saved (the output of its second layer).
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from ._lib import PackEntry
from .graph import GraphCSR
from .ops import pad16


class _MLPSlots:
    """Where one 4-layer MLP lives: flat-buffer offsets of its parameters and its packed operands."""

    def __init__(self):
        self.w_off: List[int] = []               # flat offset (floats) of the block of each weight we own
        self.b_off: List[int] = []
        self.ld: List[int] = []                  # row stride (floats) of the fp32 master matrix
        self.scale_off: Optional[int] = None
        self.shape: List[Tuple[int, int]] = []   # real (n, k) of the 4 blocks
        self.packed: List[torch.Tensor] = []     # bf16 operand views [n_pad][k_pad]
        self.bias: List[torch.Tensor] = []       # fp32 views with n_pad readable entries
        self.scale: Optional[torch.Tensor] = None


class EPDEngine:
    """`model` exposes hidden_size, processor_list (GraphNetBlocks), only_processor and, unless
    only_processor, nodes_encoder / edges_encoder / decode_module (4-layer MLP containers)."""

    def __init__(self, model: nn.Module):
        self.model = model
        self.H = int(model.hidden_size)
        if self.H not in (32, 64, 128):
            raise ValueError(f"hidden_size={self.H}: the fused kernels support 32, 64 and 128")
        self.L = len(model.processor_list)
        self.only_processor = bool(model.only_processor)
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("graphphysics_b200 runs on CUDA devices only (there is no CPU fallback)")
        # Saving h1 / h3 as well costs the forward more (two extra tile stores per MLP, each holding the
        # tile buffer until the copy engine has read it) than the backward gains from skipping the
        # recompute: 15.4 vs 14.3 ms/step on the benchmark.  Kept as an option.
        self.save_all = os.environ.get("GP_B200_SAVE_ALL", "0") == "1"
        # activation checkpointing: segments of this many processor layers (0 = keep every layer's activations).  On by
        # the reference's switch training.enable_vram_optimizations (set_memory_optimized_training), or GP_B200_CHECKPOINT=k
        from .models.layers import use_memory_optimized_training
        self.checkpoint_every = int(os.environ.get("GP_B200_CHECKPOINT", "0")) or (4 if use_memory_optimized_training() else 0)
        # the engine packs the operand copies of the weights once per step, several launches before any MLP
        # kernel reads them, so the kernels may overlap their prologue with the previous kernel's tail
        self._overlap = os.environ.get("GP_B200_NO_PDL") is None
        self._build_flat()
        self._build_packed()
        H = self.H
        # Per-CTA weight-gradient partials: kNRegions regions, one per backward launch whose reduction is
        # still pending, so one reduction launch serves a whole processor layer (5 kernels).
        stride_max = max(ops.bwd_layout(H, 128, 128)[5], 3 * H * H)
        self._region_elems = ops.sm_count() * stride_max
        # The reduction of a layer runs on a side stream under the next layer's backward kernels (their grids leave SMs
        # free; nothing on the main chain needs the reduced gradient before the optimizer): two sets of regions, so the
        # next layer writes the other set while this one is being reduced.  GP_B200_SIDE_REDUCE=0: in-stream.
        self._side_reduce = os.environ.get("GP_B200_SIDE_REDUCE", "1") != "0"
        self._side_stream = torch.cuda.Stream(device=self.device) if self._side_reduce else None
        self._side_stream2 = torch.cuda.Stream(device=self.device) if self._side_reduce else None
        self._set, self._set_events = 0, [None, None]
        self.partials_all = torch.empty(2 * self.kNRegions * self._region_elems, dtype=torch.float32, device=self.device)
        self._pending: list = []
        self._region = 0
        self._sq_ws = torch.empty(256, dtype=torch.float32, device=self.device)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=self.device)

    @staticmethod
    def _group_fusable(model: nn.Module, params):
        """Order the parameters so that the weights (and the biases) of layers that read the same input sit next to each
        other in the flat buffer: q / k / v of an Attention and linear1 / linear2 of a GatedMLP then ARE one stacked
        [3H, H] / [6H, H] matrix, and graphphysics_b200.dense runs each group as a single GEMM (forward, dgrad, wgrad)."""
        from .models.layers import Attention, GatedMLP
        groups = []
        for m in model.modules():
            if isinstance(m, Attention) and m.q_proj.weight is not m.k_proj.weight:
                lins = (m.q_proj, m.k_proj, m.v_proj)
            elif isinstance(m, GatedMLP):
                lins = (m.linear1, m.linear2)
            else:
                continue
            groups.append([l.weight for l in lins])
            if all(l.bias is not None for l in lins):
                groups.append([l.bias for l in lins])
        grouped = {id(p) for g in groups for p in g}
        if len(grouped) != sum(len(g) for g in groups):          # a parameter shared between groups: leave the order alone
            return params
        known = {id(p) for p in params}
        head = {id(g[0]): g for g in groups if all(id(p) in known for p in g)}
        out = []
        for p in params:
            if id(p) in head:
                out.extend(head[id(p)])
            elif id(p) not in grouped or not any(id(p) == id(q) for g in head.values() for q in g):
                out.append(p)
        return out

    # ------------------------------------------------------------------ parameters
    def _build_flat(self):
        named = list(self.model.named_parameters())
        offs, total = {}, 0
        for name, p in named:
            offs[name] = total
            total += (p.numel() + 15) // 16 * 16       # 64-byte slots: a short bias can be read padded
        flat = torch.zeros(total, dtype=torch.float32, device=self.device)
        for name, p in named:
            o = offs[name]
            flat[o:o + p.numel()].copy_(p.data.reshape(-1).float())
            p.data = flat[o:o + p.numel()].view(p.shape)
        self.flat = flat.requires_grad_(True)
        self.gflat = torch.zeros_like(flat)
        self.offsets: Dict[str, int] = offs
        self.shapes = {name: tuple(p.shape) for name, p in named}
        self.numels = {name: p.numel() for name, p in named}
        self.bind_param_grads()

    def bind_param_grads(self) -> None:
        """Every nn.Parameter's `.grad` is a view of the flat gradient buffer the backward kernels write,
        so torch optimizers, `clip_grad_norm_(model.parameters())` and Lightning's `configure_optimizers`
        see the gradients (the parameters themselves are views of `flat`, so in-place updates land there).
        Each backward OVERWRITES the buffer (zero_grad-then-backward semantics, what the reference's
        training_step does); re-bound after every backward because `zero_grad(set_to_none=True)` drops it."""
        for name, p in self.model.named_parameters():
            o = self.offsets[name]
            g = p.grad
            if g is None or g.data_ptr() != self.gflat.data_ptr() + 4 * o:
                p.grad = self.gflat[o:o + p.numel()].view(p.shape)

    def is_bound(self) -> bool:
        """False once somebody replaced the storage of ANY parameter (model.to(...), load_state_dict(assign=True),
        a standalone engine built on one of the blocks): the owner rebuilds the engine then."""
        base = self.flat.data_ptr()
        named = list(self.model.named_parameters())
        if len(named) != len(self.offsets):
            return False
        return all(name in self.offsets and p.data_ptr() == base + 4 * self.offsets[name] for name, p in named)

    def views_of(self, buf: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {n: buf[o:o + self.numels[n]].view(self.shapes[n]) for n, o in self.offsets.items()}

    def grads_by_name(self) -> Dict[str, torch.Tensor]:
        return self.views_of(self.gflat)

    def _build_packed(self):
        H, dev = self.H, self.device
        entries: List[PackEntry] = []
        total = [0]
        plan = []

        def alloc(n_pad, k_pad):
            off = total[0]
            total[0] += n_pad * k_pad
            return off

        def add_entry(src_name, col0, n, k, dst_off, ld_dst, row0=0):
            e = PackEntry()
            e.src_off, e.ld_src, e.src_col0 = self.offsets[src_name], self.shapes[src_name][1], col0
            e.n, e.k = n, k
            e.dst_off, e.ld_dst, e.dst_row0, e.dst_col0 = dst_off, ld_dst, row0, 0
            entries.append(e)

        def mlp_slots(prefix: str, first_cols: Optional[Tuple[int, int]] = None, norm: bool = True) -> _MLPSlots:
            s = _MLPSlots()
            for i in range(4):
                wn, bn = f"{prefix}.{2 * i}.weight", f"{prefix}.{2 * i}.bias"
                n, k = self.shapes[wn]
                col0 = 0
                if i == 0 and first_cols is not None:
                    col0, k = first_cols
                s.w_off.append(self.offsets[wn] + col0)
                s.b_off.append(self.offsets[bn])
                s.ld.append(self.shapes[wn][1])
                s.shape.append((n, k))
                n_pad, k_pad = pad16(n), pad16(k)
                dst_off = alloc(n_pad, k_pad)
                add_entry(wn, col0, n, k, dst_off, k_pad)
                plan.append((s, i, dst_off, n_pad, k_pad))
            if norm:
                s.scale_off = self.offsets[f"{prefix}.7.scale"]
            return s

        self.enc_n = self.enc_e = self.dec = None
        if not self.only_processor:
            self.enc_n = mlp_slots("nodes_encoder")
            self.enc_e = mlp_slots("edges_encoder")
            self.dec = mlp_slots("decode_module", norm=False)
        self.edge: List[_MLPSlots] = []
        self.node: List[_MLPSlots] = []
        proj_off: List[int] = []
        for l in range(self.L):
            pe, pn = f"processor_list.{l}.edge_block", f"processor_list.{l}.node_block"
            self.edge.append(mlp_slots(pe, first_cols=(0, H)))       # W1e: columns of e
            self.node.append(mlp_slots(pn, first_cols=(H, H)))       # W1a: columns of agg
            dst_off = alloc(3 * H, H)                                # Wp = [W1d ; W1s ; W1x]
            add_entry(f"{pe}.0.weight", H, H, H, dst_off, H, row0=0)
            add_entry(f"{pe}.0.weight", 2 * H, H, H, dst_off, H, row0=H)
            add_entry(f"{pn}.0.weight", 0, H, H, dst_off, H, row0=2 * H)
            proj_off.append(dst_off)

        self.packed = torch.zeros(total[0], dtype=torch.bfloat16, device=dev)
        for s, i, dst_off, n_pad, k_pad in plan:
            s.packed.append(self.packed[dst_off:dst_off + n_pad * k_pad].view(n_pad, k_pad))
            s.bias.append(self.flat.data[s.b_off[i]:s.b_off[i] + n_pad])
        for s in [self.enc_n, self.enc_e, self.dec] + self.edge + self.node:
            if s is not None and s.scale_off is not None:
                s.scale = self.flat.data[s.scale_off:s.scale_off + H]
        self.proj = [self.packed[o:o + 3 * H * H].view(3 * H, H) for o in proj_off]
        arr = (PackEntry * len(entries))(*entries)
        self.pack_table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self.n_pack = len(entries)
        self.refresh_weights()

    def refresh_weights(self):
        """fp32 masters -> packed bf16 operands (one launch).  Runs at the start of every forward."""
        ops.pack_weights(self.flat.data, self.packed, self.pack_table, self.n_pack)

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _pad_cols(t: torch.Tensor, k_pad: int) -> torch.Tensor:
        out = torch.zeros((t.shape[0], k_pad), dtype=torch.bfloat16, device=t.device)
        out[:, : t.shape[1]] = t
        return out

    def _saved(self, rows: int, save: bool):
        """Buffers for the hidden activations the backward reads back: (h1, h2, h3) bf16 [rows, H].
        h2 is needed (it is the input of backward stage B); with save_all, h1 / h3 replace the recompute
        of layers 0 / 2 -- and with it the gathers of stage A -- at the price of 4H more bytes per row."""
        if not save:
            return None
        mk = lambda: torch.empty((rows, self.H), dtype=torch.bfloat16, device=self.device)
        return (mk(), mk(), mk()) if self.save_all else (None, mk(), None)

    @staticmethod
    def _save_kw(sv):
        return {} if sv is None else dict(save_h1=sv[0], save_h2=sv[1], save_h3=sv[2])

    def _mlp(self, s: _MLPSlots, rows, a, ka, out, n_valid, **kw):
        return ops.mlp_fwd(rows, self.H, s.packed, s.bias, a=a, ka=ka, norm_scale=s.scale, out=out, n_valid=n_valid, **kw)

    def run_block(self, l: int, x: torch.Tensor, e: torch.Tensor, g: GraphCSR, bnd: torch.Tensor, save: bool):
        """One GraphNetBlock on latent bf16 x [N,H] and receiver-sorted e [E,H]."""
        H, dev, bf = self.H, self.device, torch.bfloat16
        N, E = g.num_nodes, g.num_edges
        P = torch.empty((N, 3 * H), dtype=bf, device=dev)
        ops.mlp_fwd(N, H, [self.proj[l]], [None], a=x, ka=H, out=P, n_valid=3 * H)
        e2 = torch.empty((E, H), dtype=bf, device=dev)
        agg = torch.empty((N, H), dtype=bf, device=dev)        # segment sums, rounded once where they are produced
        h2e = self._saved(E, save)
        self._mlp(self.edge[l], E, e, H, e2, H, **self._save_kw(h2e), resid=e, init=P, init_off0=0, init_off1=H,
                  idx0=g.dst, idx1=g.src, two_inits=True, seg_id=g.dst, seg_out=agg, seg_bnd=bnd, tag="edge_fwd")
        ops.seg_fixup(g.rowptr_dst, H, bnd, agg)
        x2 = torch.empty((N, H), dtype=bf, device=dev)
        h2n = self._saved(N, save)
        self._mlp(self.node[l], N, agg, H, x2, H, **self._save_kw(h2n), resid=x, init=P, init_off0=2 * H)
        return x2, e2, ((x, e, P, agg, h2e, h2n) if save else None)

    def forward(self, *args, **kwargs):
        """See _forward.  Launch overlap (programmatic dependent launch) is on only inside the engine's own
        launch sequences, where no parameter tensor is written right before an MLP kernel."""
        prev = ops.set_launch_overlap(self._overlap)
        try:
            return self._forward(*args, **kwargs)
        finally:
            ops.set_launch_overlap(prev)

    def backward(self, *args, **kwargs):
        """See _backward (launch overlap as in forward)."""
        prev = ops.set_launch_overlap(self._overlap)
        try:
            return self._backward(*args, **kwargs)
        finally:
            ops.set_launch_overlap(prev)

    def _forward(self, x_in: torch.Tensor, edge_attr: torch.Tensor, g: GraphCSR, save: bool, after_block=None):
        """x_in [N, node_in] / edge_attr [E, edge_in] fp32 in the caller's edge order (latent
        [N,H] / [E,H] when only_processor).  Returns (out, e_last_sorted, ctx).
        `after_block(x)` (optional) runs on the node latent after the encoder and after every block
        -- the halo exchange of the node-partitioned mode (dist/partitioned.py)."""
        H, dev = self.H, self.device
        N, E = g.num_nodes, g.num_edges
        bf = torch.bfloat16
        ctx = {"g": g, "layers": []} if save else None
        self.refresh_weights()
        if self.only_processor:
            x = x_in.to(bf).contiguous()
            e = edge_attr[g.perm_dst64].to(bf).contiguous()
        else:
            xin_p = self._pad_cols(x_in, self.enc_n.packed[0].shape[1])
            ea_p = self._pad_cols(edge_attr[g.perm_dst64], self.enc_e.packed[0].shape[1])
            x = torch.empty((N, H), dtype=bf, device=dev)
            e = torch.empty((E, H), dtype=bf, device=dev)
            h2n0, h2e0 = self._saved(N, save), self._saved(E, save)
            with self._beside():        # the node encoder (N rows) runs beside the edge encoder (E rows)
                self._mlp(self.enc_n, N, xin_p, xin_p.shape[1], x, H, **self._save_kw(h2n0))
            self._mlp(self.enc_e, E, ea_p, ea_p.shape[1], e, H, **self._save_kw(h2e0))
            self._rejoin()
            if save:
                ctx.update(xin_p=xin_p, ea_p=ea_p, h2n0=h2n0, h2e0=h2e0)
        bnd = torch.empty(ops.seg_bnd_size(E, H), dtype=torch.float32, device=dev)
        # memory-optimised training (the reference's enable_vram_optimizations / torch checkpointing, layers.py:24-36,
        # 803-814): keep only the inputs (x, e) of every `ckpt`-th layer; the backward re-runs the forward of one
        # segment at a time to rebuild what the layers in it need (bit-identical: every kernel is deterministic)
        ckpt = self.checkpoint_every if (save and after_block is None) else 0
        if save:
            ctx["ckpt"] = ckpt
        for l in range(self.L):
            if ckpt and l % ckpt == 0:
                ctx["layers"].append((x, e))
            x2, e2, saved = self.run_block(l, x, e, g, bnd, save and not ckpt)
            if after_block is not None:
                after_block(x2)
            if save and not ckpt:
                ctx["layers"].append(saved)
            x, e = x2, e2
        if self.only_processor:
            return x, e, ctx
        out_size = self.dec.shape[3][0]
        out = torch.empty((N, out_size), dtype=torch.float32, device=dev)
        h2d = self._saved(N, save)
        ops.mlp_fwd(N, H, self.dec.packed, self.dec.bias, a=x, ka=H, out=out, n_valid=out_size, **self._save_kw(h2d))
        if save:
            ctx.update(x_last=x, h2d=h2d)
        return out, e, ctx

    # ------------------------------------------------------------------ backward
    kNRegions = 5

    @property
    def partials(self) -> torch.Tensor:
        """Partial-gradient region of the next backward launch (the previous ones stay untouched until
        `_flush_reduce`)."""
        if self._region >= self.kNRegions:
            self._flush_reduce()
        r = self._set * self.kNRegions + self._region
        return self.partials_all[r * self._region_elems:(r + 1) * self._region_elems]

    def _queue_reduce(self, grid: int, stride: int, segs) -> None:
        """Queue the reduction of the launch that just wrote the current region, and move on to the next."""
        base = self.partials_all.data_ptr() + 4 * (self._set * self.kNRegions + self._region) * self._region_elems
        self._pending += [tuple(sg) + (base, grid, stride) for sg in segs]
        self._region += 1

    def _flush_reduce(self) -> None:
        if self._pending:
            if self._side_reduce:
                cur, side = torch.cuda.current_stream(self.device), self._side_stream
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    ops.reduce_multi(None, 0, 0, self._pending)
                ev = torch.cuda.Event()
                ev.record(side)
                self._set_events[self._set] = ev
                self._set ^= 1
                if self._set_events[self._set] is not None:      # the set written next: its last reduction must have read it
                    cur.wait_event(self._set_events[self._set])
            else:
                ops.reduce_multi(None, 0, 0, self._pending)
        self._pending, self._region = [], 0

    def _beside(self):
        """Context: launches inside go to the second side stream (after everything queued on the current stream so
        far); `_rejoin()` makes the current stream wait for them.  Without side streams: a no-op."""
        import contextlib
        if not self._side_reduce:
            return contextlib.nullcontext()
        self._side_stream2.wait_stream(torch.cuda.current_stream(self.device))
        return torch.cuda.stream(self._side_stream2)

    def _rejoin(self) -> None:
        if self._side_reduce:
            torch.cuda.current_stream(self.device).wait_stream(self._side_stream2)

    def _join_reduce(self) -> None:
        """Every queued reduction has run and the flat gradient buffer is final for the current stream."""
        self._flush_reduce()
        if self._side_reduce:
            torch.cuda.current_stream(self.device).wait_stream(self._side_stream)
            self._set_events = [None, None]

    def _reduce_stage(self, grid, ka, nb, s: _MLPSlots, ia: int, ib: int, with_scale: bool):
        """Per-CTA partial blocks of a stage over layers (ia, ib) of MLP `s` -> flat gradient buffer."""
        H = self.H
        o_dwb, o_dwa, o_dbb, o_dba, o_dsc, stride = ops.bwd_layout(H, ka, nb)
        gp = self.gflat.data_ptr()
        (nb_r, kb_r), (na_r, ka_r) = s.shape[ib], s.shape[ia]
        segs = [
            (o_dwb, nb_r, kb_r, H, gp + 4 * s.w_off[ib], s.ld[ib], False),
            (o_dwa, na_r, ka_r, ka, gp + 4 * s.w_off[ia], s.ld[ia], False),
            (o_dbb, 1, nb_r, nb, gp + 4 * s.b_off[ib], nb_r, False),
            (o_dba, 1, na_r, H, gp + 4 * s.b_off[ia], na_r, False),
        ]
        if with_scale:
            segs.append((o_dsc, 1, H, H, gp + 4 * s.scale_off, H, False))
        self._queue_reduce(grid, stride, segs)

    def _mlp_backward(self, s: _MLPSlots, rows, *, a_in, ka, h2, top, first=None, out=None, out_resid=None,
                      delta_a_out=None, seg=None, tag=None):
        """Backward through the 4-layer MLP `s`: stage B over layers (2,3), stage A over (0,1).
        top = dict(gy=..., gy_gather=..., gy_idx=...) for a normalised MLP, or dict(delta_b=...)."""
        H, dev = self.H, self.device
        h1, h2, h3 = h2 if isinstance(h2, tuple) else (None, h2, None)
        delta2 = torch.empty((rows, H), dtype=torch.bfloat16, device=dev)
        gridB = ops.mlp_bwd_stage(rows, H, a=h2, ka=H, wa=s.packed[2], ba=s.bias[2], wb=s.packed[3], bb=s.bias[3],
                                  partials=self.partials, norm_scale=s.scale, out=delta2, mask_by_ain=True,
                                  ha_saved=h3, tag=(tag + "_B") if tag else None, **top)
        self._reduce_stage(gridB, H, s.packed[3].shape[0], s, 2, 3, s.scale is not None)
        kw = dict(first or {}) if h1 is None else dict(ha_saved=h1)     # saved h1: no gathers, no recompute
        if seg is not None:
            kw.update(seg_id=seg[0], seg_out=seg[1], seg_bnd=seg[2])
        gridA = ops.mlp_bwd_stage(rows, H, a=a_in, ka=ka, wa=s.packed[0], ba=s.bias[0], wb=s.packed[1], bb=s.bias[1],
                                  partials=self.partials, delta_b=delta2, out=out, out_resid=out_resid,
                                  delta_a_out=delta_a_out, tag=(tag + "_A") if tag else None, **kw)
        self._reduce_stage(gridA, ka, H, s, 0, 1, False)

    def grad_range(self, prefix: str):
        """[lo, hi) of the flat buffers covered by the parameters whose name starts with `prefix` (contiguous: the
        buffers follow named_parameters order)."""
        names = [n for n in self.offsets if n.startswith(prefix)]
        lo = min(self.offsets[n] for n in names)
        hi = max(self.offsets[n] + self.numels[n] for n in names)
        return lo, hi

    def _backward(self, ctx, d_out: torch.Tensor, dE_sorted: Optional[torch.Tensor] = None, before_block=None, grads_ready=None):
        """Fills self.gflat with d loss / d parameters.  d_out is d loss / d output ([N,out] fp32, or
        [N,H] with only_processor); dE_sorted (bf16, receiver-sorted) is the gradient of the last
        edge latent when somebody consumes it.  Returns (dX_in, dE_in_sorted) for only_processor.
        `before_block(dX)` (optional) runs on the fp32 gradient of a block's node output before that
        block's backward -- the transpose of forward's `after_block` (reverse halo exchange).
        `grads_ready(lo, hi)` (optional) is called as soon as gflat[lo:hi] is final (decoder, then every processor
        layer from the last to the first, then the encoders): data-parallel training all-reduces that slice while the
        backward of the earlier layers is still running."""
        H, dev = self.H, self.device
        g: GraphCSR = ctx["g"]
        N, E = g.num_nodes, g.num_edges
        bf = torch.bfloat16
        if self.only_processor:
            dX = d_out.float().contiguous()
        else:
            out_size = self.dec.shape[3][0]
            Gp = torch.zeros((N, self.dec.packed[3].shape[0]), dtype=bf, device=dev)
            Gp[:, :out_size] = d_out
            dX = torch.empty((N, H), dtype=torch.float32, device=dev)
            self._mlp_backward(self.dec, N, a_in=ctx["x_last"], ka=H, h2=ctx["h2d"], top=dict(delta_b=Gp), out=dX)
            if grads_ready is not None:
                self._join_reduce()
                grads_ready(*self.grad_range("decode_module."))
        # nobody consumes the last edge latent: its gradient is zero.  An explicit zero tile keeps the top
        # layer on the same specialised kernel as the others (a memset is cheaper than the general path).
        dE = dE_sorted if dE_sorted is not None else torch.zeros((E, H), dtype=bf, device=dev)
        bnd = torch.empty(ops.seg_bnd_size(E, H, backward=True), dtype=torch.float32, device=dev)
        gp = self.gflat.data_ptr()
        ckpt = ctx.get("ckpt", 0)
        seg_saved, fwd_bnd = {}, None
        for l in reversed(range(self.L)):
            if ckpt:
                if l not in seg_saved:                    # rebuild the saved activations of the segment that holds layer l
                    seg_saved.clear()
                    s0 = (l // ckpt) * ckpt
                    xs, es = ctx["layers"][s0 // ckpt]
                    if fwd_bnd is None:
                        fwd_bnd = torch.empty(ops.seg_bnd_size(E, H), dtype=torch.float32, device=dev)
                    for ll in range(s0, l + 1):
                        xs, es, seg_saved[ll] = self.run_block(ll, xs, es, g, fwd_bnd, True)
                x, e, P, agg, h2e, h2n = seg_saved.pop(l)
            else:
                x, e, P, agg, h2e, h2n = ctx["layers"][l]
            if before_block is not None:
                before_block(dX)
            # node MLP:  x' = x + norm(MLP([x, agg]))
            dagg = torch.empty((N, H), dtype=bf, device=dev)       # rounded where it is produced (its consumer stages it as bf16 anyway)
            dQ = torch.empty((N, H), dtype=bf, device=dev)
            self._mlp_backward(self.node[l], N, a_in=agg, ka=H, h2=h2n, top=dict(gy=dX), out=dagg, delta_a_out=dQ,
                               first=dict(init=P, init_off0=2 * H))
            # edge MLP:  e' = e + u,  agg = segment-sum(u)   =>   du = dE' + dagg[dst]
            dE_new = torch.empty((E, H), dtype=bf, device=dev)
            d1 = torch.empty((E, H), dtype=bf, device=dev)
            dPd = torch.empty((N, H), dtype=bf, device=dev)
            self._mlp_backward(self.edge[l], E, a_in=e, ka=H, h2=h2e, top=dict(gy=dE, gy_gather=dagg, gy_idx=g.dst),
                               out=dE_new, out_resid=dE, delta_a_out=d1, seg=(g.dst, dPd, bnd),
                               first=dict(init=P, init_off0=0, init_off1=H, idx0=g.dst, idx1=g.src, two_inits=True),
                               tag="edge_bwd")
            dPs = torch.empty((N, H), dtype=bf, device=dev)
            with self._beside():      # the receiver-side fix-up (a few CTAs) runs beside the sender-side segment sum
                ops.seg_fixup(g.rowptr_dst, H, bnd, dPd, backward=True)
            ops.segsum_gather(d1, g.perm_src, g.rowptr_src, H, dPs)
            self._rejoin()            # both feed linear_bwd
            # projection P = x . Wp^T, plus the residual path of x
            dX_new = torch.empty((N, H), dtype=torch.float32, device=dev)
            grid = ops.linear_bwd(N, H, [dPd, dPs, dQ], self.proj[l], x, dX, dX_new, self.partials)
            pe, pn = f"processor_list.{l}.edge_block.0.weight", f"processor_list.{l}.node_block.0.weight"
            self._queue_reduce(grid, 3 * H * H, [
                (0, H, H, H, gp + 4 * (self.offsets[pe] + H), 3 * H, False),
                (H * H, H, H, H, gp + 4 * (self.offsets[pe] + 2 * H), 3 * H, False),
                (2 * H * H, H, H, H, gp + 4 * self.offsets[pn], 2 * H, False),
            ])
            self._flush_reduce()      # one reduction launch per processor layer
            if grads_ready is not None:
                self._join_reduce()
                grads_ready(*self.grad_range(f"processor_list.{l}."))
            dX, dE = dX_new, dE_new
        if self.only_processor:
            self._join_reduce()
            return dX, dE
        with self._beside():
            self._mlp_backward(self.enc_n, N, a_in=ctx["xin_p"], ka=ctx["xin_p"].shape[1], h2=ctx["h2n0"], top=dict(gy=dX))
        self._mlp_backward(self.enc_e, E, a_in=ctx["ea_p"], ka=ctx["ea_p"].shape[1], h2=ctx["h2e0"], top=dict(gy=dE))
        self._rejoin()
        self._join_reduce()
        if grads_ready is not None:
            lo_n, hi_n = self.grad_range("nodes_encoder.")
            lo_e, hi_e = self.grad_range("edges_encoder.")
            grads_ready(min(lo_n, lo_e), max(hi_n, hi_e))
        return None, None

    # ------------------------------------------------------------------ optimizer helpers
    def grad_sqnorm(self) -> torch.Tensor:
        ops.sqnorm(self.gflat, self._sq_ws, self.sqnorm)
        return self.sqnorm


def _adjacent(*ts) -> bool:
    """True when the tensors are consecutive slices of one buffer (so they read as one stacked matrix / vector)."""
    return all(b.data_ptr() == a.data_ptr() + a.numel() * a.element_size() and
               b.untyped_storage().data_ptr() == a.untyped_storage().data_ptr() for a, b in zip(ts, ts[1:]))


class FlatParams:
    """Flat fp32 parameter / gradient buffers for a model that runs under torch autograd (the
    Transformer path): every nn.Parameter becomes a view of `flat`, every `.grad` a view of `gflat`,
    so clip + AdamW run as the same two kernels (gp_sqnorm, gp_adamw) as on the EPD engine."""

    def __init__(self, model: nn.Module):
        params = [p for p in model.parameters() if p.requires_grad]
        uniq, seen = [], set()
        for p in params:                      # shared weights (use_separate_proj_weight=False) appear once
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        uniq = self._group_fusable(model, uniq)
        self.device = uniq[0].device
        total = sum((p.numel() + 15) // 16 * 16 for p in uniq)
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.gflat = torch.zeros_like(self.flat)
        off = 0
        for p in uniq:
            n = p.numel()
            self.flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + n].view(p.shape)
            p.grad = self.gflat[off:off + n].view(p.shape)
            off += (n + 15) // 16 * 16
        self._sq_ws = torch.empty(256, dtype=torch.float32, device=self.device)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=self.device)

    @staticmethod
    def _group_fusable(model: nn.Module, params):
        """Order the parameters so that the weights (and the biases) of layers that read the same input sit next to each
        other in the flat buffer: q / k / v of an Attention and linear1 / linear2 of a GatedMLP then ARE one stacked
        [3H, H] / [6H, H] matrix, and graphphysics_b200.dense runs each group as a single GEMM (forward, dgrad, wgrad)."""
        from .models.layers import Attention, GatedMLP
        groups = []
        for m in model.modules():
            if isinstance(m, Attention) and m.q_proj.weight is not m.k_proj.weight:
                lins = (m.q_proj, m.k_proj, m.v_proj)
            elif isinstance(m, GatedMLP):
                lins = (m.linear1, m.linear2)
            else:
                continue
            groups.append([l.weight for l in lins])
            if all(l.bias is not None for l in lins):
                groups.append([l.bias for l in lins])
        grouped = {id(p) for g in groups for p in g}
        if len(grouped) != sum(len(g) for g in groups):          # a parameter shared between groups: leave the order alone
            return params
        known = {id(p) for p in params}
        head = {id(g[0]): g for g in groups if all(id(p) in known for p in g)}
        out = []
        for p in params:
            if id(p) in head:
                out.extend(head[id(p)])
            elif id(p) not in grouped or not any(id(p) == id(q) for g in head.values() for q in g):
                out.append(p)
        return out

    def zero_grad(self):
        self.gflat.zero_()

    def grad_sqnorm(self) -> torch.Tensor:
        ops.sqnorm(self.gflat, self._sq_ws, self.sqnorm)
        return self.sqnorm


class EPDFunction(torch.autograd.Function):
    """The whole encode-process-decode model as one autograd node over the flat parameter buffer."""

    @staticmethod
    def forward(ctx, flat, x_in, edge_attr, engine: EPDEngine, g: GraphCSR):
        out, _, saved = engine.forward(x_in, edge_attr, g, save=True)
        ctx.engine, ctx.saved = engine, saved
        return out

    @staticmethod
    def backward(ctx, d_out):
        eng = ctx.engine
        eng.backward(ctx.saved, d_out.contiguous())
        ctx.saved = None
        eng.bind_param_grads()
        gflat = eng.gflat
        if eng.flat.grad is not None and eng.flat.grad.data_ptr() == gflat.data_ptr():
            gflat = gflat.clone()        # autograd would otherwise add the buffer to itself
        return gflat, None, None, None, None


class BlockFunction(torch.autograd.Function):
    """Processor-only engine (a stack of GraphNetBlocks) with gradients for the latent inputs too."""

    @staticmethod
    def forward(ctx, flat, x, edge_attr, engine: EPDEngine, g: GraphCSR):
        x_out, e_sorted, saved = engine.forward(x, edge_attr, g, save=True)
        ctx.engine, ctx.saved, ctx.g = engine, saved, g
        e_out = torch.empty_like(e_sorted)
        e_out[g.perm_dst64] = e_sorted
        return x_out.float(), e_out.float()

    @staticmethod
    def backward(ctx, dx, de):
        eng, g = ctx.engine, ctx.g
        de_sorted = de[g.perm_dst64].to(torch.bfloat16).contiguous()
        dX, dE = eng.backward(ctx.saved, dx.contiguous(), de_sorted)
        ctx.saved = None
        eng.bind_param_grads()
        ge = torch.empty((g.num_edges, eng.H), dtype=torch.float32, device=dx.device)
        ge[g.perm_dst64] = dE.float()
        gflat = eng.gflat
        if eng.flat.grad is not None and eng.flat.grad.data_ptr() == gflat.data_ptr():
            gflat = gflat.clone()
        return gflat, dX, ge, None, None
