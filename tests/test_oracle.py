"""The oracle (our C restatement) against the reference: committed golden streams, and the
compiled reference itself when oracle/_ref is present.  CPU only."""
import hashlib

import numpy as np
import pytest

from conftest import golden_stream, load_manifest
from golden_inputs import make_input, raw_bytes
from oracle import oracle as O
from zipnn_b200 import ZipNN

CASES = load_manifest()


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


@pytest.mark.parametrize("rec", CASES, ids=[c["name"] for c in CASES])
def test_port_reproduces_reference_stream(rec):
    data = make_input(rec["input"])
    raw = raw_bytes(data)
    if sha(raw) != rec["input_sha256"]:
        pytest.skip("input generator drifted on this machine (numpy/torch RNG or cast)")
    z = ZipNN(**rec["ctor"])
    plan = z.plan(data if rec["ctor"]["input_format"] == "torch" else raw)
    assert plan["header"][:24].hex() == rec["header_hex"][:48]
    stream = O.zipnn_compress(plan["header"], np.frombuffer(raw, dtype=np.uint8), plan["num_buf"], plan["bit_reorder"],
                              plan["byte_reorder"], plan["chunk"], plan["threshold"], threads=4).tobytes()
    assert len(stream) == rec["stream_len"]
    assert sha(stream) == rec["stream_sha256"]
    gold = golden_stream(rec)
    if gold is not None:
        assert stream == gold
    body = np.frombuffer(stream, dtype=np.uint8)[len(plan["header"]):]
    back = O.zipnn_decompress(body, plan["num_buf"], plan["bit_reorder"], plan["byte_reorder"], plan["chunk"], len(raw), threads=4)
    assert back.tobytes() == raw


def test_port_matches_compiled_reference_blocks():
    if O.ref_cdll() is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(3)
    for it in range(300):
        size = int(rng.choice([12, 13, 64, 257, 1500, 4096, 65536, 131072, int(rng.integers(1, 131073))]))
        kind = it % 5
        if kind == 0:
            x = (rng.standard_normal(size) * 0.02).astype(np.float32)
            src = (x.view(np.uint32) >> 23).astype(np.uint8)
        elif kind == 1:
            src = rng.integers(0, 256, size, dtype=np.uint8)
        elif kind == 2:
            k = int(rng.integers(2, 256))
            src = rng.choice(k, size, p=rng.dirichlet(np.ones(k) * rng.uniform(0.01, 1))).astype(np.uint8)
        elif kind == 3:
            src = np.full(size, 7, dtype=np.uint8)
        else:
            src = np.minimum(rng.geometric(rng.uniform(0.02, 0.9), size), 255).astype(np.uint8)
        cap = 256 * 1024 if it % 2 else 128 * 1024
        assert O.huf_compress(src, cap) == O.ref_huf_compress(src, cap)


def test_port_matches_compiled_reference_streams():
    ref = O.ref_core()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(11)
    for it in range(40):
        G = [1, 2, 4][it % 3]
        bits = (it // 3) % 2
        bm = 220 if G == 4 else 10
        chunk = 128 * 1024 if G == 1 else int(rng.choice([256 * 1024, 65536, 4096]))
        nelem = int(rng.choice([1, 3, 13, 4096, chunk // G + 1, int(rng.integers(1, 200000))]))
        n = nelem * G
        if it % 4 == 3:
            data = rng.integers(0, 256, n, dtype=np.uint8)
        else:
            x = (rng.standard_normal(max(n // 2, 1) + 2) * 0.02).astype(np.float32)
            data = np.ascontiguousarray((x.view(np.uint32) >> 16).astype(np.uint16).view(np.uint8)[:n])
        h = bytearray(32)
        h[0:2] = b"ZN"
        r = bytes(ref.zipnn_core(bytes(h), bytearray(data.tobytes()), G, bits, bm, 0, chunk, 0.95, 10, 4))
        o = O.zipnn_compress(h, data, G, bits, bm, chunk, 0.95, threads=2).tobytes()
        assert r == o
        assert bytes(ref.combine_dtype(o[32:], G, bits, bm, chunk, n, 2)) == data.tobytes()
