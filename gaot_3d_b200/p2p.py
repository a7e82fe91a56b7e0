"""Collectives of the intra-sample sharded path over NVLink peer memory (csrc/a2a.cu) instead of NCCL kernels.

The sharded step's exchanges are small and many: 40 all-to-alls of ~3 MB inside the transformer, a reduce-scatter and an
all-gather of the 17 MB latent field each way, one 44 MB gradient all-reduce.  As NCCL collectives they cost 24-29 us,
150-390 us and 560 us apiece at 8 ranks (profiles/r02d_trace_shard8m_8_rank*.txt) -- launch latency, protocol hand-shakes and
the skew they absorb, not bandwidth.  Here every rank owns SYMMETRIC buffers (torch.distributed._symmetric_memory: cuMem
allocations mapped into all ranks of the node) and the exchanges are plain kernels on the compute stream:
    all-to-all      one kernel stores block j straight into slot `rank` of peer j's buffer, then the allocation's barrier
    all-gather      the same kernel with the one block stored into every peer
    reduce-scatter  local copy into the own buffer, barrier, one kernel PULLS the own slab from every peer and sums it in
                    rank order (deterministic, unlike a ring whose order depends on the rank)
    all-reduce      two-shot: reduce-scatter into the own slab, all-gather of the reduced slabs
Each kind alternates between TWO buffers: a peer's stores of exchange k+2 land in the buffer of exchange k only after that
peer has passed the barrier of exchange k+1, which this rank joins only after (in stream order) its consumers of exchange k.
Everything is capturable in CUDA graphs (kernels + the barrier kernel); buffers are sized on first use, OUTSIDE capture.
If the peer mapping cannot be set up (no P2P access, no handle exchange in the container) or GAOT_A2A=nccl is set, every
function answers None and the callers keep their torch.distributed (NCCL / gloo) collectives; the choice is all-reduced so
that all ranks agree.  New with the sharded path: the reference has sample-level DDP only
(src/trainer/stat.py:431-436).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch
import torch.distributed as dist

_CTX = {}        # id(group) -> _Peer | False
_MIN_BYTES = {"a2a": 16 << 20, "gather": 64 << 20, "reduce": 64 << 20}
# which exchanges take the peer path (GAOT_P2P_OPS, comma separated): a2a, ag (all-gather), rs (reduce-scatter), ar (all-reduce)
_OPS = set(os.environ.get("GAOT_P2P_OPS", "a2a,ag,rs,ar").split(","))


class _Peer:
    def __init__(self, group, dev):
        self.group, self.dev = group, dev
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.pools = {}          # kind -> dict(bufs, hdls, ptrs, bytes, turn)

    def pool(self, kind: str, nbytes: int):
        """Two symmetric buffers of >= nbytes for this kind of exchange (collective on first use / growth; never under capture)."""
        p = self.pools.get(kind)
        if p is not None and p["bytes"] >= nbytes:
            return p
        if torch.cuda.is_current_stream_capturing():
            return None
        import torch.distributed._symmetric_memory as symm
        g = dist.group.WORLD if self.group is None else self.group
        # generous first allocation (the step's largest exchange of each kind fits: 44 MB gradient all-reduce, 17 MB latent
        # field, 13 MB [q|k|v]) so that the pools do not grow in steady state; if one must grow, every rank first drains its GPU
        # and meets the others: a peer may still be pulling from the buffer that is about to be released
        size = max(int(nbytes), _MIN_BYTES.get(kind, 1 << 20))
        size = (size + 255) // 256 * 256
        if p is not None:
            torch.cuda.synchronize(self.dev)
            dist.barrier(group=self.group)
        bufs, hdls, ptrs = [], [], []
        for _ in range(2):
            b = symm.empty(size, dtype=torch.uint8, device=self.dev)
            h = symm.rendezvous(b, group=g.group_name)
            bufs.append(b)
            hdls.append(h)
            ptrs.append((ctypes.c_void_p * self.world)(*[int(x) for x in h.buffer_ptrs]))
        p = dict(bufs=bufs, hdls=hdls, ptrs=ptrs, bytes=size, turn=0)
        self.pools[kind] = p
        return p

    @staticmethod
    def take(p):
        t = p["turn"]
        p["turn"] = t ^ 1
        return p["bufs"][t], p["hdls"][t], p["ptrs"][t]


def _peer(group, dev) -> Optional[_Peer]:
    key = id(group)
    ctx = _CTX.get(key)
    if ctx is not None:
        return ctx or None
    if torch.cuda.is_current_stream_capturing():
        return None
    ok = os.environ.get("GAOT_A2A", "p2p") != "nccl" and dev.type == "cuda" and dist.get_backend(group) == "nccl"
    peer = None
    if ok:
        try:
            peer = _Peer(group, dev)
            ok = peer.world <= 16 and peer.pool("probe", 1 << 20) is not None
        except Exception as e:  # no peer access / no handle exchange in this container
            ok = False
            _CTX["error"] = repr(e)
    flag = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)          # every rank takes the same path
    ok = bool(flag.item())
    _CTX[key] = peer if ok else False
    return peer if ok else None


def backend(group=None) -> str:
    return "p2p" if _CTX.get(id(group)) else "nccl"


def reset() -> None:
    _CTX.clear()


def _lib():
    from . import ops
    return ops._lib_()


def _launch(fn, *args):
    from ._lib import check
    check(fn(*args), fn.__name__ if hasattr(fn, "__name__") else "p2p")


def all_to_all(send: torch.Tensor, group=None) -> Optional[torch.Tensor]:
    """[R, ...] -> [R, ...]: block j goes to rank j, block i of the result came from rank i.  The result is a VIEW of a
    symmetric buffer, valid until the second next all-to-all: consume or copy it first (tblock.py does)."""
    from . import ops
    R = send.shape[0]
    nbytes = send.numel() * send.element_size()
    if "a2a" not in _OPS or not send.is_cuda or R == 0 or (nbytes // R) % 16 != 0:
        return None
    peer = _peer(group, send.device)
    if peer is None or R != peer.world:
        return None
    p = peer.pool("a2a", nbytes)
    if p is None:
        return None
    buf, hdl, ptrs = peer.take(p)
    send = send.contiguous()
    _launch(_lib().gaot_p2p_put, ops._p(send), ptrs, peer.rank, R, nbytes // R, 0, ops._stream(send.device))
    hdl.barrier(channel=0)
    return buf[:nbytes].view(send.dtype).view(send.shape)


def all_gather(inp: torch.Tensor, group=None) -> Optional[torch.Tensor]:
    """concat over ranks of `inp` (row slabs) as a fresh tensor [R * m, ...]."""
    from . import ops
    nbytes = inp.numel() * inp.element_size()
    if "ag" not in _OPS or not inp.is_cuda or nbytes == 0 or nbytes % 16 != 0:
        return None
    peer = _peer(group, inp.device)
    if peer is None:
        return None
    p = peer.pool("gather", nbytes * peer.world)
    if p is None:
        return None
    buf, hdl, ptrs = peer.take(p)
    inp = inp.contiguous()
    _launch(_lib().gaot_p2p_put, ops._p(inp), ptrs, peer.rank, peer.world, nbytes, 1, ops._stream(inp.device))
    hdl.barrier(channel=0)
    return buf[:nbytes * peer.world].view(inp.dtype).view((inp.shape[0] * peer.world,) + tuple(inp.shape[1:])).clone()


def _reduce_slab_from(peer: _Peer, p, src: torch.Tensor, slab_floats: int, out: torch.Tensor):
    """copy `src` (fp32, flat) into the own symmetric buffer, barrier, pull-and-sum this rank's slab from every peer"""
    from . import ops
    buf, hdl, ptrs = peer.take(p)
    n = src.numel()
    buf[:n * 4].view(torch.float32).copy_(src.reshape(-1))
    hdl.barrier(channel=0)
    _launch(_lib().gaot_p2p_reduce, ptrs, peer.world, peer.rank * slab_floats, slab_floats, ops._p(out), ops._stream(src.device))


def reduce_scatter(inp: torch.Tensor, group=None) -> Optional[torch.Tensor]:
    """(sum over ranks of inp)[slab of this rank] for fp32 inp [M, ...] with M divisible by the ranks; fixed summation order."""
    if "rs" not in _OPS or not inp.is_cuda or inp.dtype != torch.float32 or inp.numel() == 0:
        return None
    peer = _peer(group, inp.device)
    if peer is None:
        return None
    R = peer.world
    if inp.shape[0] % R or (inp.numel() // R) % 4:
        return None
    p = peer.pool("reduce", inp.numel() * 4)
    if p is None:
        return None
    slab = inp.numel() // R
    out = torch.empty((inp.shape[0] // R,) + tuple(inp.shape[1:]), dtype=torch.float32, device=inp.device)
    _reduce_slab_from(peer, p, inp.contiguous(), slab, out)
    return out


def all_reduce_(flat: torch.Tensor, group=None) -> bool:
    """In-place SUM over the ranks of a flat fp32 tensor (two-shot: pull-reduce of the own slab, all-gather of the slabs).
    Returns False (and leaves `flat` alone) when the peer path is unavailable."""
    from . import ops
    if "ar" not in _OPS or not flat.is_cuda or flat.dtype != torch.float32 or flat.dim() != 1 or flat.numel() == 0:
        return False
    peer = _peer(group, flat.device)
    if peer is None:
        return False
    R = peer.world
    n = flat.numel()
    slab = (n + R - 1) // R
    slab = (slab + 3) // 4 * 4                                   # every slab a multiple of 16 bytes; the tail is padding
    p_red = peer.pool("reduce", slab * R * 4)
    p_gat = peer.pool("gather", slab * R * 4)
    if p_red is None or p_gat is None:
        return False
    buf, hdl, ptrs = peer.take(p_red)
    stage = buf[:slab * R * 4].view(torch.float32)
    stage[:n].copy_(flat)
    if slab * R > n:
        stage[n:].zero_()
    hdl.barrier(channel=0)
    mine = torch.empty(slab, dtype=torch.float32, device=flat.device)
    _launch(_lib().gaot_p2p_reduce, ptrs, R, peer.rank * slab, slab, ops._p(mine), ops._stream(flat.device))
    gbuf, ghdl, gptrs = peer.take(p_gat)
    _launch(_lib().gaot_p2p_put, ops._p(mine), gptrs, peer.rank, R, slab * 4, 1, ops._stream(flat.device))
    ghdl.barrier(channel=0)
    flat.copy_(gbuf[:slab * R * 4].view(torch.float32)[:n])
    return True
