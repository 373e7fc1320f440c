"""Compare windows of a (large) ZipNN stream with a checker's stream for the same chunks.

A chunk range [c0, c1) of a stream is self-describing once its rows of the two tables are
rebased (reference layout, csrc/zipnn_core.c:105-244: `[header][types u8[G][K]][cum u64le[G][K]]
[payload group-major]`, cum inclusive and per group, u64 -- :145-153).  So the checker (the
compiled reference or the oracle port; the CALLER supplies it, this module imports neither)
only has to compress the input bytes of the window, and every byte of the big stream that
belongs to the window -- type bytes, size-table rows and the per-group payload slices -- can
be compared with it.  That is how a 16 GiB stream, whose cumulative offsets pass 2^32, is
pinned to the reference without running the CPU codec over all of it.
"""
from __future__ import annotations

import numpy as np


def _np(x, a: int, b: int) -> np.ndarray:
    """bytes [a, b) of a stream held in a torch tensor (CPU/CUDA) or a numpy array."""
    if isinstance(x, np.ndarray):
        return x[a:b]
    return x[a:b].cpu().numpy()


class StreamTables:
    """Type bytes and cumulative sizes of a whole stream (host copies: 9*G*K bytes)."""

    def __init__(self, stream, hdr_len: int, G: int, K: int):
        self.stream, self.hdr_len, self.G, self.K = stream, hdr_len, G, K
        self.types = _np(stream, hdr_len, hdr_len + G * K).reshape(G, K).copy()
        raw = _np(stream, hdr_len + G * K, hdr_len + 9 * G * K)
        self.cum = np.frombuffer(raw.tobytes(), dtype="<u8").reshape(G, K).astype(np.uint64)
        self.payload0 = hdr_len + 9 * G * K
        tot = self.cum[:, -1].astype(np.uint64)
        self.base = self.payload0 + np.concatenate([[0], np.cumsum(tot)[:-1]]).astype(np.uint64)
        self.total = int(self.payload0 + int(tot.sum()))

    def first_chunk_past(self, g: int, limit: int) -> int:
        """first chunk whose inclusive cumulative size in group g exceeds `limit` (K if none)."""
        return int(np.searchsorted(self.cum[g], np.uint64(limit), side="right"))


def compare_window(tab: StreamTables, c0: int, c1: int, win_stream: np.ndarray, win_hdr_len: int) -> int:
    """win_stream = the checker's stream for input chunks [c0, c1) alone (any header of
    win_hdr_len bytes).  Raises AssertionError on the first difference; returns the number
    of bytes of the big stream that were compared."""
    G = tab.G
    Kw = c1 - c0
    w = np.asarray(win_stream, dtype=np.uint8)
    wt = w[win_hdr_len: win_hdr_len + G * Kw].reshape(G, Kw)
    wc = np.frombuffer(w[win_hdr_len + G * Kw: win_hdr_len + 9 * G * Kw].tobytes(), dtype="<u8").reshape(G, Kw)
    assert np.array_equal(tab.types[:, c0:c1], wt), f"type bytes differ in chunks [{c0},{c1})"
    lo = tab.cum[:, c0 - 1] if c0 else np.zeros(G, dtype=np.uint64)
    assert np.array_equal(tab.cum[:, c0:c1] - lo.reshape(G, 1), wc), f"cumulative sizes differ in chunks [{c0},{c1})"
    compared = 9 * G * Kw
    off = win_hdr_len + 9 * G * Kw
    for g in range(G):
        ln = int(wc[g, -1])
        a = int(tab.base[g]) + int(lo[g])
        got = _np(tab.stream, a, a + ln)
        assert got.size == ln and np.array_equal(got, w[off: off + ln]), f"group {g} payload differs in chunks [{c0},{c1})"
        off += ln
        compared += ln
    assert off == w.size, "checker stream has trailing bytes"
    return compared


def check_stream_windows(stream, hdr_len: int, G: int, K: int, chunk: int, n: int, input_bytes, windows, compress_window):
    """stream: the big stream; input_bytes(a, b) -> numpy uint8 of the input bytes [a, b);
    windows: [(c0, c1)]; compress_window(np.uint8 array) -> (checker stream, its header length).
    -> dict(bytes_compared, windows, max_offset) (raises on any difference)."""
    tab = StreamTables(stream, hdr_len, G, K)
    total_len = stream.numel() if hasattr(stream, "numel") else stream.size
    assert tab.total == total_len, "stream length does not match its size table"
    done = 0
    seen = []
    top = 0
    for c0, c1 in windows:
        c0, c1 = max(0, c0), min(K, c1)
        if c1 <= c0:
            continue
        data = input_bytes(c0 * chunk, min(n, c1 * chunk))
        ws, whl = compress_window(data)
        done += compare_window(tab, c0, c1, np.asarray(ws, dtype=np.uint8), whl)
        seen.append([c0, c1])
        top = max(top, int(tab.base[G - 1]) + int(tab.cum[G - 1, c1 - 1]))
    return {"bytes_compared": int(done), "windows": seen, "max_stream_offset": int(top)}
