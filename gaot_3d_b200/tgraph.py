"""The latent processor (GAOT3D.process: patchify -> patch_linear -> PE -> Transformer -> un-patchify, reference
src/model/gaot_3d.py:166-222 and src/model/layers/attn.py:298-325) as TWO CUDA graphs: one for the forward, one for the backward.

Why: at S = 16384 the 10-block transformer is ~450 launches of this library's kernels per step; issued one by one from
Python they leave ~4 ms of gaps in a 28 ms step on one GPU and make the intra-sample sharded step host-bound at 8 ranks
(DESIGN.md r01 section 6).  The shapes of this part never change from sample to sample (fixed latent grid), so the whole
launch sequence -- including the all-to-alls of the sequence-parallel mode, NCCL being capturable -- is recorded once per
(model, shape, parallel mode) and replayed with one host call each way.

How: `ProcessorRunner` drives the SAME per-block code the eager path uses (tblock._BlockFn.forward / .backward,
ops._LinearFn) with a stand-in for the autograd ctx, so graph and eager paths cannot diverge; activations saved for the
backward live in the graphs' private memory pool.  `_GraphedProcessFn` is the single autograd node the model sees.

Sequence-parallel slab mode (shard.py): the runner works on this rank's D-slab of the latent grid (contiguous latent-token
range -> contiguous patch range, because tokens and patches are both D-major) with tblock's "sp" mode inside the blocks.

Falls back to the eager module path whenever something is outside this envelope (dropout active, a second forward before the
backward of the first, unsupported shapes, GAOT_NO_GRAPH=1).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn

from . import ops, tblock
from .ops import _LinearFn

_STATE = {"enabled": os.environ.get("GAOT_NO_GRAPH", "0") != "1", "runners": {}}


def set_enabled(on: bool) -> None:
    """Process-wide switch (bench.py turns it off for its per-kernel timing pass: events cannot be read inside a replay)."""
    _STATE["enabled"] = bool(on)


def enabled() -> bool:
    return _STATE["enabled"]


def reset() -> None:
    _STATE["runners"].clear()


class _Ctx:
    """Stand-in for torch.autograd's ctx so that the Function bodies can be driven by hand (and recorded into a graph)."""

    def __init__(self, needs_input_grad):
        self.needs_input_grad = needs_input_grad
        self.saved_tensors = ()

    def save_for_backward(self, *ts):
        self.saved_tensors = ts


def _blocks_of(proc):
    enc = list(proc.encoder_layers)
    mid = [proc.middle_layer] if proc.middle_layer is not None else []
    dec = list(proc.decoder_layers)
    return enc, mid, dec


def _block_params(blk, use_skip):
    a, f = blk.attn, blk.ffn
    return [blk.skip_proj.weight if use_skip else None, blk.skip_proj.bias if use_skip else None, blk.attn_norm.weight,
            a.q_proj.weight, a.k_proj.weight, a.v_proj.weight, a.o_proj.weight, blk.ffn_norm.weight, f.w1.weight, f.w2.weight,
            f.w3.weight]


class ProcessorRunner:
    def __init__(self, model, B: int, m_local: int, slab, par, rope: bool):
        self.model, self.B, self.m_local, self.slab, self.par, self.rope = model, int(B), int(m_local), slab, par, rope
        proc = model.processor
        self.enc, self.mid, self.dec = _blocks_of(proc)
        self.skip_on = bool(proc.use_long_range_skip)
        self.params = self._param_list()
        self.dev = model.patch_linear.weight.device
        P = model.patch_size
        r, R = slab
        self.nd_l, self.nh, self.nw, self.P, self.C = model.D // P // R, model.H // P, model.W // P, P, model.node_latent_size
        assert self.nd_l * P * model.H * model.W == self.m_local, "latent slab does not match the token count"
        self.S_l = self.nd_l * self.nh * self.nw
        self.width = P ** 3 * self.C
        self.pe = None
        if model.positional_embedding_name == "absolute":
            pos = model.positions.to(self.dev)
            self.pe = model._compute_absolute_embeddings(pos, self.width)[r * self.S_l:(r + 1) * self.S_l].contiguous()
        self.graph_f = self.graph_b = None
        self.pending = False           # a forward whose backward has not run yet owns the saved activations
        self.n_launch_f = self.n_launch_b = 0
        self.grads = {}

    # ------------------------------------------------------------------ parameters (order = order of returned gradients)
    def _param_list(self):
        m = self.model
        ps = [m.patch_linear.weight, m.patch_linear.bias]
        proc = m.processor
        for lin in (proc.input_proj, proc.output_proj):
            if not isinstance(lin, nn.Identity):
                ps += [lin.weight, lin.bias]
        for blk in self.enc + self.mid:
            ps += [p for p in _block_params(blk, False) if p is not None]
        for blk in self.dec:
            ps += [p for p in _block_params(blk, self.skip_on and blk.skip_connection) if p is not None]
        return ps

    def key_ptrs(self):
        return tuple(p.data_ptr() for p in self.params)

    # ------------------------------------------------------------------ the computation (eager or under capture)
    def _linear_f(self, lin, x):
        ctx = _Ctx((True, False, True, lin.bias is not None, False))
        y = _LinearFn.forward(ctx, x, None, lin.weight, lin.bias, None)
        self.tape.append(("lin", ctx, lin))
        return y

    def _block_f(self, blk, x, skip):
        a = blk.attn
        use_skip = skip is not None
        freqs = a.rotary_emb.freqs if (self.rope and hasattr(a, "rotary_emb")) else None
        cfg = (int(a.num_heads), int(a.num_kv_heads), float(blk.attn_norm.eps), freqs, 0.0, 0, self.par)
        ctx = _Ctx((True,) * 14)
        y = tblock._BlockFn.forward(ctx, x, skip, cfg, *_block_params(blk, use_skip))
        self.tape.append(("blk", ctx, blk, use_skip))
        return y

    def _fwd(self, x):
        B, P, C = self.B, self.P, self.C
        self.tape = []
        xp = x.view(B, self.nd_l, P, self.nh, P, self.nw, P, C).permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(B, self.S_l, self.width)
        h = self._linear_f(self.model.patch_linear, xp)
        if self.pe is not None:
            h = h + self.pe
        proc = self.model.processor
        if not isinstance(proc.input_proj, nn.Identity):
            h = self._linear_f(proc.input_proj, h)
        skips = []
        for blk in self.enc:
            h = self._block_f(blk, h, None)
            skips.append(h)
        for blk in self.mid:
            h = self._block_f(blk, h, None)
        for blk in self.dec:
            skip = skips.pop() if self.skip_on else None
            h = self._block_f(blk, h, skip if blk.skip_connection else None)
        if not isinstance(proc.output_proj, nn.Identity):
            h = self._linear_f(proc.output_proj, h)
        y = h.view(B, self.nd_l, self.nh, self.nw, P, P, P, C).permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous()
        return y.view(B, self.m_local, C)

    def _bwd(self, dy):
        B, P, C = self.B, self.P, self.C
        grads = {}
        d = dy.view(B, self.nd_l, P, self.nh, P, self.nw, P, C).permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(B, self.S_l, -1)
        n_enc = len(self.enc)
        dskips = {}
        enc_seen = 0
        dec_seen = 0
        for entry in reversed(self.tape):
            if entry[0] == "lin":
                _, ctx, lin = entry
                dx, _, dw, db, _ = _LinearFn.backward(ctx, d)
                grads[lin.weight] = dw
                if lin.bias is not None:
                    grads[lin.bias] = db
                d = dx
                continue
            _, ctx, blk, use_skip = entry
            is_enc = any(blk is e for e in self.enc)
            if is_enc:
                j = n_enc - 1 - enc_seen                   # encoder blocks come off the tape last-first
                enc_seen += 1
                if j in dskips:
                    d = d + dskips.pop(j)
            out = tblock._BlockFn.backward(ctx, d)
            dx, dskip, _, dwsk, dbsk, dn1, dwq, dwk, dwv, dwo, dn2, dw1, dw2, dw3 = out
            a, f = blk.attn, blk.ffn
            for p, g in ((blk.attn_norm.weight, dn1), (a.q_proj.weight, dwq), (a.k_proj.weight, dwk), (a.v_proj.weight, dwv),
                         (a.o_proj.weight, dwo), (blk.ffn_norm.weight, dn2), (f.w1.weight, dw1), (f.w2.weight, dw2), (f.w3.weight, dw3)):
                grads[p] = g
            if use_skip:
                grads[blk.skip_proj.weight] = dwsk
                if blk.skip_proj.bias is not None:
                    grads[blk.skip_proj.bias] = dbsk
                # decoder block i (tape order reversed: last decoder first) consumed the output of encoder block n-1-i
                i = len(self.dec) - 1 - dec_seen
                dskips[n_enc - 1 - i] = dskip.reshape(d.shape)
            if any(blk is e for e in self.dec):
                dec_seen += 1
            d = dx.reshape(B, self.S_l, -1)
        self.grads = grads
        dx = d.view(B, self.nd_l, self.nh, self.nw, P, P, P, C).permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous()
        return dx.view(B, self.m_local, C)

    # ------------------------------------------------------------------ capture / replay
    def capture(self):
        lib = ops._lib_()
        dev = self.dev
        self.static_x = torch.zeros(self.B, self.m_local, self.C, device=dev)
        self.static_dy = torch.zeros(self.B, self.m_local, self.C, device=dev)
        # warm-up on a side stream: lazy CUDA state (function attributes, NCCL communicators) must exist before capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._fwd(self.static_x)
            self._bwd(self.static_dy)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.tape, self.grads = [], {}
        pool = torch.cuda.graph_pool_handle()
        self.graph_f, self.graph_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        n0 = int(lib.gaot_launch_count())
        with torch.cuda.graph(self.graph_f, pool=pool, capture_error_mode="thread_local"):
            self.static_y = self._fwd(self.static_x)
        n1 = int(lib.gaot_launch_count())
        with torch.cuda.graph(self.graph_b, pool=pool, capture_error_mode="thread_local"):
            self.static_dx = self._bwd(self.static_dy)
        n2 = int(lib.gaot_launch_count())
        self.n_launch_f, self.n_launch_b = n1 - n0, n2 - n1
        lib.gaot_launch_count_add(-(n2 - n0))           # recording is not launching
        self.static_grads = [self.grads[p] for p in self.params]

    def forward(self, x):
        self.static_x.copy_(x.reshape(self.static_x.shape))
        self.graph_f.replay()
        ops._lib_().gaot_launch_count_add(self.n_launch_f)
        return self.static_y.clone()

    def backward(self, dy):
        self.static_dy.copy_(dy.reshape(self.static_dy.shape))
        self.graph_b.replay()
        ops._lib_().gaot_launch_count_add(self.n_launch_b)
        return self.static_dx.clone(), [g.clone() for g in self.static_grads]


class _GraphedProcessFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, runner, *params):
        ctx.runner = runner
        runner.pending = True
        return runner.forward(x)

    @staticmethod
    def backward(ctx, dy):
        runner = ctx.runner
        dx, grads = runner.backward(dy.contiguous())
        runner.pending = False
        return (dx, None, *grads)


def _supported(model, x, par) -> bool:
    from .layers.attn import FUSED_BLOCK
    if not (enabled() and FUSED_BLOCK and x.is_cuda and x.dtype == torch.float32):
        return False
    proc = model.processor
    enc, mid, dec = _blocks_of(proc)
    blocks = enc + mid + dec
    if not blocks:
        return False
    probe = torch.empty(0, device=x.device)
    for blk in blocks:
        if blk.training and blk.attn.atten_dropout > 0.0:
            return False                                  # a captured dropout seed would repeat the same mask every step
        hid = blk.attn.q_proj.in_features
        if not blk._fused_ok(probe.new_empty(1, 1, hid), None):
            return False
    lins = [model.patch_linear] + [l for l in (proc.input_proj, proc.output_proj) if not isinstance(l, nn.Identity)]
    return all(ops.linear_supported(l.in_features, l.out_features) for l in lins)


def process(model, rndata: torch.Tensor, slab=(0, 1)) -> Optional[torch.Tensor]:
    """Graph-replayed GAOT3D.process on `rndata` [B, m_local, C] (m_local = this rank's latent slab; the whole grid when
    slab = (0, 1)).  Returns None when the configuration is outside the envelope (caller uses the eager module path).
    The parallel mode of the blocks is tblock's process-wide one (tblock.set_head_parallel), as on the eager path."""
    enc, mid, dec = _blocks_of(model.processor)
    blocks = enc + mid + dec
    par = tblock._head_shard(int(blocks[0].attn.num_heads), int(blocks[0].attn.num_kv_heads)) if blocks else None
    if not _supported(model, rndata, par):
        return None
    B, m_local, _ = rndata.shape
    rope = model.positional_embedding_name == "rope"
    need_grad = torch.is_grad_enabled() and (rndata.requires_grad or any(p.requires_grad for p in model.processor.parameters()))
    key = (id(model), B, m_local, tuple(slab), None if par is None else (par[0], par[1], par[3]), rope)
    runner = _STATE["runners"].get(key)
    if runner is not None and runner.key_ptrs() != runner.ptrs_at_capture:
        runner = None                                      # parameters were re-allocated (.to(), load with assign): re-record
    if runner is None:
        runner = ProcessorRunner(model, B, m_local, slab, par, rope)
        runner.capture()
        runner.ptrs_at_capture = runner.key_ptrs()
        while len(_STATE["runners"]) >= 4:                 # a handful of live (model, shape) pairs at most: drop the oldest
            _STATE["runners"].pop(next(iter(_STATE["runners"])))
        _STATE["runners"][key] = runner
    if runner.pending and need_grad:
        return None                                        # two live forwards: the second one runs eagerly
    x = rndata.contiguous()
    if not need_grad:
        with torch.no_grad():
            return runner.forward(x)
    return _GraphedProcessFn.apply(x, runner, *runner.params)
