"""Chunk-range sharding of one ZipNN stream across the GPUs of a node.

The reference has no distributed code (SURVEY.md sections 2.2, 5).  Chunks are independent units
-- own histogram, table and bitstreams (reference csrc/zipnn_core.c:294-388, 768-858) -- and
the only cross-chunk state in the stream is, per byte group, the inclusive prefix sum of
payload sizes and the group-major payload order (csrc/zipnn_core.c:105-244).  So:

  compress    every rank codes its own contiguous chunk range with no communication; an
              all-gather of G payload totals per rank (a few dozen bytes) fixes every offset;
              the optional `gather_stream` then moves each rank's per-group payload once, with
              point-to-point sends straight into its final position in the owner's buffer
              (NCCL over NVLink on GPUs), producing the byte-identical single-GPU stream.
  decompress  the owner sends each rank the metadata rows and the per-group payload byte
              ranges of its chunk range; every rank rebuilds a self-contained local stream
              and decodes its shard.  Nothing is exchanged after decode.

One process per GPU; `torch.distributed` is plumbing (backend nccl on GPUs, gloo in the CPU
tests).  The codec calls are injectable so the host-side logic is testable without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

HEADER_LEN = 32


def chunk_range(K: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous chunk range owned by `rank`: ceil(K/world) chunks each, the tail may be short or empty."""
    per = (K + world - 1) // world
    return min(K, rank * per), min(K, (rank + 1) * per)


def byte_range(n: int, chunk: int, rank: int, world: int) -> Tuple[int, int]:
    K = (n + chunk - 1) // chunk
    c0, c1 = chunk_range(K, rank, world)
    return min(n, c0 * chunk), min(n, c1 * chunk)


@dataclass
class StreamMeta:
    """Positions inside one stream (python header included)."""
    hdr_len: int
    G: int
    K: int
    types: np.ndarray      # [G, K] uint8
    cum: np.ndarray        # [G, K] uint64 (inclusive, per group)
    payload0: int          # offset of group 0's payload
    base: List[int]        # payload offset of group g, relative to payload0

    @property
    def total(self) -> int:
        return self.payload0 + (int(self.cum[:, -1].sum()) if self.K else 0)


def parse_meta(stream_u8: np.ndarray, hdr_len: int, G: int, K: int) -> StreamMeta:
    t0 = hdr_len
    types = np.array(stream_u8[t0: t0 + G * K], dtype=np.uint8).reshape(G, K)
    cum = np.frombuffer(np.array(stream_u8[t0 + G * K: t0 + 9 * G * K], dtype=np.uint8).tobytes(), dtype="<u8").reshape(G, K).copy()
    base, acc = [], 0
    for g in range(G):
        base.append(acc)
        acc += int(cum[g, -1]) if K else 0
    return StreamMeta(hdr_len, G, K, types, cum, hdr_len + 9 * G * K, base)


def _u8(t: torch.Tensor) -> torch.Tensor:
    t = t.detach().contiguous().reshape(-1)
    return t if t.dtype == torch.uint8 else t.view(torch.uint8)


def _bytes_tensor(b: bytes, device) -> torch.Tensor:
    return torch.frombuffer(bytearray(b), dtype=torch.uint8).to(device)


def _ipc_alias(t: Optional[torch.Tensor], src: int, group=None) -> torch.Tensor:
    """Every rank gets a tensor that ALIASES rank `src`'s CUDA tensor `t` (CUDA IPC: the ranks are processes on
    one node).  Copies to / from it are peer copies over NVLink, issued by the rank that owns the other side --
    no send/recv pairing, no staging through NCCL's channel buffers."""
    from torch.multiprocessing.reductions import reduce_tensor
    rank = dist.get_rank(group)
    box = [reduce_tensor(t) if rank == src else None]
    dist.broadcast_object_list(box, src, group=group)
    if rank == src:
        return t
    fn, args = box[0]
    return fn(*args)


def _use_ipc(t: torch.Tensor, transport: str) -> bool:
    return transport == "ipc" or (transport == "auto" and t.is_cuda and dist.get_backend() == "nccl")


def _p2p(ops):
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


# ----------------------------------------------------------------------------------------------
def gather_stream(local_stream: torch.Tensor, local_hdr_len: int, G: int, K_local: int, n_local: int,
                  global_header: bytes, dst: int = 0, group=None, chunk: Optional[int] = None,
                  transport: str = "auto", cache: Optional[dict] = None) -> Optional[torch.Tensor]:
    """Merge per-rank streams (each covering a contiguous chunk range, in rank order) into the
    single stream the reference would have produced for the concatenated input.

    `global_header` is the python-level header (+ packed shape) of the whole tensor; bytes
    [16:24] (original length) and [24:32] (stream length) are filled in here.
    Returns the stream on `dst`, None elsewhere.
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = local_stream.device
    ls = _u8(local_stream)
    meta_host = ls[: local_hdr_len + 9 * G * K_local].cpu().numpy()
    m = parse_meta(meta_host, local_hdr_len, G, K_local)
    mine = torch.tensor([K_local, n_local] + [int(m.cum[g, -1]) if K_local else 0 for g in range(G)], dtype=torch.int64, device=dev)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)          # the only collective: (2 + G) integers per rank
    allv = [v.cpu().tolist() for v in allv]
    Ks = [v[0] for v in allv]
    ns = [v[1] for v in allv]
    # Shards must be cut at chunk boundaries (`byte_range`): a short chunk in the middle would give a stream the
    # reference cannot decode.  Only the last non-empty rank may end with a ragged chunk; no rank may follow an empty one.
    nonempty = [r for r in range(world) if ns[r]]
    if nonempty and nonempty != list(range(len(nonempty))):
        raise ValueError("sharded compress: an empty shard precedes a non-empty one; use zipnn_b200.sharded.byte_range to cut shards")
    if chunk:
        for r in nonempty[:-1]:
            if ns[r] % chunk or Ks[r] * chunk != ns[r]:
                raise ValueError(f"sharded compress: rank {r}'s shard ({ns[r]} bytes) is not a whole number of {chunk}-byte chunks; "
                                 "use zipnn_b200.sharded.byte_range to cut shards")
    tot = [[v[2 + g] for g in range(G)] for v in allv]   # tot[r][g]
    K = sum(Ks)
    H = len(global_header)
    group_total = [sum(tot[r][g] for r in range(world)) for g in range(G)]
    payload0 = H + 9 * G * K
    group_base = [payload0 + sum(group_total[:g]) for g in range(G)]
    total = payload0 + sum(group_total)
    c_off = [sum(Ks[:r]) for r in range(world)]
    rank_base = [[sum(tot[q][g] for q in range(r)) for g in range(G)] for r in range(world)]  # inside group g

    # rows this rank contributes: types[g][local chunks], cum[g][local chunks] rebased to the group
    types_rows = torch.from_numpy(m.types.copy()).to(dev)                      # [G, K_local]
    cum_rows = m.cum.astype(np.int64) + np.array(rank_base[rank], dtype=np.int64).reshape(G, 1)
    cum_rows = torch.from_numpy(cum_rows).to(dev)

    if _use_ipc(ls, transport):
        # ---- peer copies: every rank writes its rows and payload slices straight into the owner's buffer ----
        # The owner's buffer and the other ranks' aliases of it are kept between calls (`cache`): sharing a fresh
        # allocation costs a pickled handle broadcast and an IPC open per rank, a few ms each.
        if cache is not None and cache.get("cap", 0) >= total and cache.get("dst") == dst:
            view = cache["view"]
        else:
            cap = total + (total >> 3) if cache is not None else total
            buf = torch.empty(cap, dtype=torch.uint8, device=dev) if rank == dst else None
            view = _ipc_alias(buf, dst, group)
            if cache is not None:
                cache.clear()
                cache.update(cap=cap, dst=dst, view=view)
        out = view[:total] if rank == dst else None
        if rank == dst:
            hdr = bytearray(global_header)
            hdr[16:24] = int(sum(ns)).to_bytes(8, "little")
            hdr[24:32] = int(total).to_bytes(8, "little")
            view[:H] = _bytes_tensor(bytes(hdr), dev)
        if K_local:
            c0 = c_off[rank]
            cum_bytes = cum_rows.contiguous().view(torch.uint8).view(G, 8 * K_local)
            for g in range(G):
                view[H + g * K + c0: H + g * K + c0 + K_local].copy_(types_rows[g], non_blocking=True)
                a = H + G * K + 8 * (g * K + c0)
                view[a: a + 8 * K_local].copy_(cum_bytes[g], non_blocking=True)
                if tot[rank][g]:
                    a = group_base[g] + rank_base[rank][g]
                    src0 = m.payload0 + m.base[g]
                    view[a: a + tot[rank][g]].copy_(ls[src0: src0 + tot[rank][g]], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        if rank != dst and cache is None:
            del view
        dist.barrier(group=group)
        return out
    if rank == dst:
        out = torch.empty(total, dtype=torch.uint8, device=dev)
        hdr = bytearray(global_header)
        hdr[16:24] = int(sum(ns)).to_bytes(8, "little")
        hdr[24:32] = int(total).to_bytes(8, "little")
        out[:H] = _bytes_tensor(bytes(hdr), dev)
        types_all = out[H: H + G * K].view(G, K)
        cum_all = torch.empty((G, K), dtype=torch.int64, device=dev)
        ops, staged = [], []
        for r in range(world):
            if Ks[r] == 0:
                continue
            sl = slice(c_off[r], c_off[r] + Ks[r])
            if r == dst:
                types_all[:, sl] = types_rows
                cum_all[:, sl] = cum_rows
                for g in range(G):
                    a = group_base[g] + rank_base[r][g]
                    src0 = m.payload0 + m.base[g]
                    out[a: a + tot[r][g]] = ls[src0: src0 + tot[r][g]]
            else:
                tb = torch.empty((G, Ks[r]), dtype=torch.uint8, device=dev)
                cb = torch.empty((G, Ks[r]), dtype=torch.int64, device=dev)
                staged.append((sl, tb, cb))
                ops.append(dist.P2POp(dist.irecv, tb, r, group))
                ops.append(dist.P2POp(dist.irecv, cb, r, group))
                for g in range(G):
                    if tot[r][g]:
                        a = group_base[g] + rank_base[r][g]
                        ops.append(dist.P2POp(dist.irecv, out[a: a + tot[r][g]], r, group))  # lands in place
        _p2p(ops)
        for sl, tb, cb in staged:
            types_all[:, sl] = tb
            cum_all[:, sl] = cb
        out[H + G * K: H + 9 * G * K] = cum_all.reshape(-1).view(torch.uint8)
        return out
    if K_local:
        ops = [dist.P2POp(dist.isend, types_rows.contiguous(), dst, group), dist.P2POp(dist.isend, cum_rows.contiguous(), dst, group)]
        for g in range(G):
            if tot[rank][g]:
                src0 = m.payload0 + m.base[g]
                ops.append(dist.P2POp(dist.isend, ls[src0: src0 + tot[rank][g]], dst, group))
        _p2p(ops)
    return None


def scatter_stream(stream: Optional[torch.Tensor], hdr_len: int, G: int, chunk: int, src: int = 0, group=None,
                   device=None, transport: str = "auto") -> Tuple[torch.Tensor, int, bytes]:
    """Inverse of gather_stream.  The owner (`src`) holds the whole stream; every rank gets back
    (local_stream, local_orig_bytes, global_header) where local_stream is a self-contained
    stream -- 32-byte header, its rows of the tables, its payload -- for its chunk range."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rank == src:
        s = _u8(stream)
        dev = s.device
        head = [s.numel(), hdr_len]
    else:
        dev = torch.device(device) if device is not None else torch.device("cpu")
        head = [0, 0]
    hv = torch.tensor(head, dtype=torch.int64, device=dev)
    dist.broadcast(hv, src, group=group)
    total, hdr_len = [int(x) for x in hv.cpu().tolist()]
    hdr_t = s[:hdr_len].clone() if rank == src else torch.empty(hdr_len, dtype=torch.uint8, device=dev)
    dist.broadcast(hdr_t, src, group=group)
    gheader = hdr_t.cpu().numpy().tobytes()
    n = int.from_bytes(gheader[16:24], "little")
    K = (n + chunk - 1) // chunk
    meta_t = s[hdr_len: hdr_len + 9 * G * K].clone() if rank == src else torch.empty(9 * G * K, dtype=torch.uint8, device=dev)
    dist.broadcast(meta_t, src, group=group)            # ~9 bytes per (group, chunk): 1.2 MB for 16 GiB of bf16
    meta = parse_meta(np.concatenate([np.zeros(hdr_len, np.uint8), meta_t.cpu().numpy()]), hdr_len, G, K)

    def span(r, g):
        c0, c1 = chunk_range(K, r, world)
        if c1 <= c0:
            return 0, 0
        lo = int(meta.cum[g, c0 - 1]) if c0 else 0
        return meta.payload0 + meta.base[g] + lo, int(meta.cum[g, c1 - 1]) - lo

    c0, c1 = chunk_range(K, rank, world)
    Kl = c1 - c0
    b0, b1 = byte_range(n, chunk, rank, world)
    n_local = b1 - b0
    sizes = [span(rank, g)[1] for g in range(G)]
    local_total = HEADER_LEN + 9 * G * Kl + sum(sizes)
    local = torch.empty(local_total, dtype=torch.uint8, device=dev)
    lh = bytearray(gheader[:HEADER_LEN])
    lh[16:24] = int(n_local).to_bytes(8, "little")
    lh[24:32] = int(local_total).to_bytes(8, "little")
    local[:HEADER_LEN] = _bytes_tensor(bytes(lh), dev)
    if Kl:
        local[HEADER_LEN: HEADER_LEN + G * Kl] = torch.from_numpy(np.ascontiguousarray(meta.types[:, c0:c1]).reshape(-1)).to(dev)
        rebased = meta.cum[:, c0:c1].astype(np.int64) - np.array([int(meta.cum[g, c0 - 1]) if c0 else 0 for g in range(G)], dtype=np.int64).reshape(G, 1)
        local[HEADER_LEN + G * Kl: HEADER_LEN + 9 * G * Kl] = torch.from_numpy(np.ascontiguousarray(rebased).reshape(-1)).to(dev).view(torch.uint8)
    if _use_ipc(local, transport):
        # ---- peer copies: every rank pulls its payload ranges out of the owner's stream ----
        view = _ipc_alias(s if rank == src else None, src, group)
        at = HEADER_LEN + 9 * G * Kl
        for g in range(G):
            off, ln = span(rank, g)
            if ln:
                local[at: at + ln].copy_(view[off: off + ln], non_blocking=True)
            at += ln
        torch.cuda.current_stream(dev).synchronize()
        if rank != src:
            del view
        dist.barrier(group=group)
        return local, n_local, gheader
    # payload: point-to-point from the owner, each piece straight into place
    ops = []
    at = HEADER_LEN + 9 * G * Kl
    if rank == src:
        for r in range(world):
            for g in range(G):
                off, ln = span(r, g)
                if not ln:
                    continue
                if r == src:
                    pos = HEADER_LEN + 9 * G * Kl + sum(sizes[:g])
                    local[pos: pos + ln] = s[off: off + ln]
                else:
                    ops.append(dist.P2POp(dist.isend, s[off: off + ln], r, group))
    else:
        for g in range(G):
            if sizes[g]:
                ops.append(dist.P2POp(dist.irecv, local[at: at + sizes[g]], src, group))
            at += sizes[g]
    _p2p(ops)
    return local, n_local, gheader


# ----------------------------------------------------------------------------------------------
class ShardedZipNN:
    """Data-parallel front end: each rank holds the slice of the flat tensor given by
    `byte_range(n, chunk, rank, world)`.

        z = ShardedZipNN()                      # after dist.init_process_group("nccl")
        stream = z.compress(local_shard)        # whole stream on rank 0 (byte-identical to 1 GPU)
        shard  = z.decompress(stream)           # every rank gets its slice back
    """

    def __init__(self, group=None, compress_local: Callable = None, decompress_local: Callable = None, **zipnn_kwargs):
        from .zipnn import ZipNN
        self.group = group
        self.kw = dict(zipnn_kwargs)
        self.kw.setdefault("input_format", "torch")
        self._ZipNN = ZipNN
        self._compress_local = compress_local
        self._decompress_local = decompress_local
        self._gather_cache = {}   # owner buffer + aliases, reused while it is large enough (the returned stream is a
                                  # view of it: valid until the next compress / gather of this object)

    def _codec(self):
        if self._compress_local is None:
            from .zipnn import _compress_device, _decompress_device
            self._compress_local, self._decompress_local = _compress_device, _decompress_device
        return self._compress_local, self._decompress_local

    def compress_local(self, local: torch.Tensor):
        """Rank-local part only: (local_stream, plan).  No communication."""
        z = self._ZipNN(**self.kw)
        plan = z.plan(local)
        comp, _ = self._codec()
        flat = _u8(local)
        hdr = bytearray(plan["header"][:HEADER_LEN])
        stream = comp(flat, bytes(hdr), plan["num_buf"], plan["bit_reorder"], plan["byte_reorder"], plan["chunk"], plan["threshold"])
        return stream, plan

    def compress(self, local: torch.Tensor, global_shape=None, dst: int = 0) -> Optional[torch.Tensor]:
        from .util_torch import zipnn_pack_shape
        stream, plan = self.compress_local(local)
        n_local = local.numel() * local.element_size()
        K_local = (n_local + plan["chunk"] - 1) // plan["chunk"]
        gh = bytearray(plan["header"][:HEADER_LEN])
        world = dist.get_world_size(self.group)
        if global_shape is None:
            cnt = torch.tensor([local.numel()], dtype=torch.int64, device=stream.device)
            dist.all_reduce(cnt, group=self.group)
            global_shape = (int(cnt.item()),)
        assert world >= 1
        return self.gather(stream, plan, n_local, global_shape, dst)

    def gather(self, local_stream: torch.Tensor, plan: dict, n_local: int, global_shape, dst: int = 0) -> Optional[torch.Tensor]:
        """The exchange half of `compress`: per-rank streams -> the single-GPU stream on `dst`."""
        from .util_torch import zipnn_pack_shape
        K_local = (n_local + plan["chunk"] - 1) // plan["chunk"]
        gh = bytearray(plan["header"][:HEADER_LEN])
        ext = zipnn_pack_shape(tuple(global_shape)) if self.kw.get("input_format", "torch") != "byte" else b""
        return gather_stream(local_stream, HEADER_LEN, plan["num_buf"], K_local, n_local, bytes(gh) + ext, dst, self.group,
                             chunk=plan["chunk"], cache=self._gather_cache)

    def decompress(self, stream: Optional[torch.Tensor], src: int = 0, device=None) -> torch.Tensor:
        from .util_torch import torch_dtype_of_code, zipnn_unpack_shape
        from .zipnn import HUF_MAX_BLOCK
        rank = dist.get_rank(self.group)
        # layout facts every rank needs before the scatter: header length, groups, chunk
        if rank == src:
            s = _u8(stream)
            head = s[: HEADER_LEN + 1 + 9 * 255].cpu().numpy().tobytes()
            code = head[15]
            G = 1 if code in (29, 30) else 4 if code in (1, 2) else 2
            chunk = 2 ** head[14]
            chunk = min(chunk, HUF_MAX_BLOCK) if G == 1 else chunk
            hdr_len = HEADER_LEN + (zipnn_unpack_shape(head[HEADER_LEN:])[1] if head[8] in (2, 3) else 0)
            facts = [G, chunk, hdr_len, head[6], head[5], code]
            dev = s.device
        else:
            facts = [0] * 6
            dev = torch.device(device) if device is not None else torch.device("cpu")
        fv = torch.tensor(facts, dtype=torch.int64, device=dev)
        dist.broadcast(fv, src, group=self.group)
        G, chunk, hdr_len, bits, bytes_mode, code = [int(x) for x in fv.cpu().tolist()]
        local, n_local, _ = scatter_stream(stream if rank == src else None, hdr_len, G, chunk, src, self.group, device=dev)
        tdt = torch_dtype_of_code(code)
        if n_local == 0:
            return torch.empty(0, dtype=tdt if tdt is not None else torch.uint8, device=local.device)
        _, dec = self._codec()
        out = dec(local[HEADER_LEN:], G, bits, bytes_mode, chunk, n_local)
        return out.view(tdt) if tdt is not None else out
