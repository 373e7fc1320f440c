"""Parity of the CUDA path against the oracle and the reference's golden streams.
Every test here calls through the C ABI (via zipnn_b200.ZipNN or ctypes directly)."""
import ctypes as C
import hashlib

import numpy as np
import pytest
import torch

from conftest import golden_stream, load_manifest
from golden_inputs import make_input, raw_bytes
from oracle import oracle as O
from zipnn_b200 import ZipNN, _native

pytestmark = pytest.mark.gpu
CASES = load_manifest()


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def _u8(t):
    return t.detach().contiguous().reshape(-1).view(torch.uint8)


# ------------------------------------------------------------------ golden streams
@pytest.mark.parametrize("rec", CASES, ids=[c["name"] for c in CASES])
def test_golden_compress_and_decompress(rec):
    data = make_input(rec["input"])
    raw = raw_bytes(data)
    if sha(raw) != rec["input_sha256"]:
        pytest.skip("input generator drifted on this machine")
    torch_fmt = rec["ctor"]["input_format"] == "torch"
    # --- device-resident: CUDA tensor in, CUDA stream out
    z = ZipNN(**rec["ctor"])
    dev_in = data.cuda() if torch_fmt else torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
    keep = _u8(dev_in).clone()
    stream = z.compress(dev_in)
    assert stream.is_cuda and stream.dtype == torch.uint8
    assert torch.equal(_u8(dev_in), keep), "compress must not modify its input"
    sbytes = stream.cpu().numpy().tobytes()
    assert len(sbytes) == rec["stream_len"]
    assert sha(sbytes) == rec["stream_sha256"]
    gold = golden_stream(rec)
    if gold is not None:
        assert sbytes == gold
    back = ZipNN(**rec["ctor"]).decompress(stream)
    assert back.is_cuda
    if torch_fmt:
        assert back.dtype == data.dtype and tuple(back.shape) == tuple(data.shape)
    assert _u8(back).cpu().numpy().tobytes() == raw
    # --- host buffers: the reference's own calling convention
    zh = ZipNN(**rec["ctor"])
    hstream = zh.compress(data.clone() if torch_fmt else raw)
    assert bytes(hstream) == sbytes
    hback = ZipNN(**rec["ctor"]).decompress(bytes(hstream))
    if torch_fmt:
        assert not hback.is_cuda and hback.dtype == data.dtype and tuple(hback.shape) == tuple(data.shape)
        assert raw_bytes(hback) == raw
    else:
        assert bytes(hback) == raw


def test_decodes_committed_reference_streams_directly():
    """Decode bytes the REFERENCE wrote (tests/golden/*.znn), not bytes we produced."""
    n = 0
    for rec in CASES:
        gold = golden_stream(rec)
        if gold is None:
            continue
        data = make_input(rec["input"])
        raw = raw_bytes(data)
        if sha(raw) != rec["input_sha256"]:
            continue
        out = ZipNN(**rec["ctor"]).decompress(torch.frombuffer(bytearray(gold), dtype=torch.uint8).cuda())
        assert _u8(out).cpu().numpy().tobytes() == raw
        n += 1
    assert n >= 10


# ------------------------------------------------------------------ oracle on seeded inputs
def _plane_inputs(rng, kind, n):
    if kind == "gauss16":
        x = (rng.standard_normal(n // 2 + 1) * 0.02).astype(np.float32)
        return np.ascontiguousarray((x.view(np.uint32) >> 16).astype(np.uint16).view(np.uint8)[:n])
    if kind == "gauss32":
        x = (rng.standard_normal(n // 4 + 1) * 0.02).astype(np.float32)
        return np.ascontiguousarray(x.view(np.uint8)[:n])
    if kind == "bytes":
        return rng.integers(0, 256, n, dtype=np.uint8)
    if kind == "zeros":
        return np.zeros(n, dtype=np.uint8)
    if kind == "few":
        return rng.choice(6, n, p=[.5, .25, .12, .06, .04, .03]).astype(np.uint8)
    if kind == "skew":
        k = int(rng.integers(130, 256))
        return (255 - rng.choice(k, n, p=rng.dirichlet(np.ones(k) * 0.3))).astype(np.uint8)
    raise ValueError(kind)


@pytest.mark.parametrize("G,bits", [(1, 0), (2, 1), (2, 0), (4, 1), (4, 0)])
def test_cabi_matches_oracle_on_seeded_inputs(G, bits):
    rng = np.random.default_rng(100 + 10 * G + bits)
    L = _native.lib()
    bm = 220 if G == 4 else 10
    kinds = ["gauss16", "gauss32", "bytes", "zeros", "few", "skew"]
    for it in range(36):
        chunk = 131072 if G == 1 else int(rng.choice([262144, 262144, 65536, 4096, 1024]))
        sizes = [G, 2 * G, 12 * G, 13 * G, chunk - G, chunk, chunk + G, 3 * chunk + 5 * G, 64, 4096,
                 int(rng.integers(1, 40000)) * G, int(rng.integers(1, 6 * chunk // G + 2)) * G]
        n = int(sizes[it % len(sizes)])
        data = _plane_inputs(rng, kinds[it % len(kinds)], n)
        hdr = bytearray(32 + (it % 5))
        hdr[0:2] = b"ZN"
        want = O.zipnn_compress(hdr, data, G, bits, bm, chunk, 0.95, threads=4)
        d_in = torch.from_numpy(data.copy()).cuda()
        bound = _native.compress_bound(n, G, chunk, len(hdr))
        d_out = torch.zeros(bound, dtype=torch.uint8, device="cuda")
        ws = torch.empty(_native.compress_workspace_size(n, G, chunk), dtype=torch.uint8, device="cuda")
        out_len = C.c_size_t(0)
        hbuf = (C.c_char * len(hdr)).from_buffer_copy(bytes(hdr))
        st = L.zipnn_b200_compress(d_in.data_ptr(), n, hbuf, len(hdr), G, bits, bm, chunk, 0.95, d_out.data_ptr(), bound,
                                   C.byref(out_len), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        assert st == 0, st
        got = d_out[: out_len.value].cpu().numpy()
        assert out_len.value == want.size, (it, n, chunk, out_len.value, want.size)
        assert np.array_equal(got, want), (it, n, chunk)
        assert np.array_equal(d_in.cpu().numpy(), data)
        # decode the ORACLE's stream with the CUDA path
        body = torch.from_numpy(want[len(hdr):].copy()).cuda()
        d_dec = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
        ws2 = torch.empty(_native.decompress_workspace_size(n, G, chunk), dtype=torch.uint8, device="cuda")
        st = L.zipnn_b200_decompress(body.data_ptr(), body.numel(), G, bits, bm, chunk, n, d_dec.data_ptr(), ws2.data_ptr(),
                                     ws2.numel(), torch.cuda.current_stream().cuda_stream, 1)
        assert st == 0, (st, it, n, chunk)
        assert np.array_equal(d_dec[:n].cpu().numpy(), data), (it, n, chunk)
        assert int(d_dec[n:].sum()) == 0, "wrote past the end of the output"


@pytest.mark.parametrize("G,bits", [(1, 0), (2, 1), (2, 0), (4, 1)])
def test_split_regroup_match_oracle(G, bits):
    rng = np.random.default_rng(7 + G)
    L = _native.lib()
    for n in [G, 16 * G, 16 * G + G, 4096 * G, 100003 * G, 262144]:
        data = rng.integers(0, 256, n, dtype=np.uint8)
        want = O.split_chunk(data, G, bits)
        stride = ((n + G - 1) // G + 15) // 16 * 16
        d_in = torch.from_numpy(data.copy()).cuda()
        planes = torch.zeros(G * stride, dtype=torch.uint8, device="cuda")
        assert L.zipnn_b200_split(d_in.data_ptr(), n, G, bits, planes.data_ptr(), stride, torch.cuda.current_stream().cuda_stream) == 0
        host = planes.cpu().numpy()
        for g in range(G):
            assert np.array_equal(host[g * stride: g * stride + want[g].size], want[g]), (n, g)
        back = torch.zeros(n, dtype=torch.uint8, device="cuda")
        assert L.zipnn_b200_regroup(planes.data_ptr(), stride, n, G, bits, back.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
        assert np.array_equal(back.cpu().numpy(), data)
        assert np.array_equal(d_in.cpu().numpy(), data)


# ------------------------------------------------------------------ edge cases + errors
def test_empty_and_tiny_tensors():
    for dt in (torch.bfloat16, torch.float32, torch.float16, torch.float8_e4m3fn):
        for n in (0, 1, 2, 3, 7):
            t = (torch.randn(n) * 0.02).to(dt)
            z = ZipNN(input_format="torch")
            s = z.compress(t.cuda())
            plan = ZipNN(input_format="torch").plan(t)
            want = O.zipnn_compress(plan["header"], np.frombuffer(raw_bytes(t), dtype=np.uint8), plan["num_buf"], plan["bit_reorder"],
                                    plan["byte_reorder"], plan["chunk"], plan["threshold"])
            assert s.cpu().numpy().tobytes() == want.tobytes(), (dt, n)
            back = ZipNN(input_format="torch").decompress(s)
            assert back.dtype == dt and back.numel() == n
            assert raw_bytes(back.cpu()) == raw_bytes(t)


def test_unaligned_views_and_noncontiguous_inputs():
    base = (torch.randn(300001) * 0.02).to(torch.bfloat16).cuda()
    for t in (base[1:], base[3:70001], base.reshape(-1)[::2]):
        s = ZipNN(input_format="torch").compress(t)
        back = ZipNN(input_format="torch").decompress(s)
        assert torch.equal(back.view(torch.int16), t.contiguous().view(torch.int16))
    # stream held at an odd offset inside a larger buffer
    s = ZipNN(input_format="torch").compress(base)
    big = torch.zeros(s.numel() + 13, dtype=torch.uint8, device="cuda")
    big[5: 5 + s.numel()] = s
    back = ZipNN(input_format="torch").decompress(big[5: 5 + s.numel()])
    assert torch.equal(back.view(torch.int16), base.view(torch.int16))


def _expect_corrupt(stream):
    """Must be rejected as a corrupt stream -- not fail with a CUDA error that merely happens to
    be a RuntimeError too -- and must leave the device usable."""
    with pytest.raises(RuntimeError, match="corrupt"):
        ZipNN(input_format="torch").decompress(stream)
    torch.cuda.synchronize()


def test_corrupt_streams_are_rejected():
    t = (torch.randn(200000) * 0.02).to(torch.bfloat16).cuda()
    s = ZipNN(input_format="torch").compress(t).clone()
    with pytest.raises(ValueError):
        bad = s.clone()
        bad[0] = 0x41
        ZipNN(input_format="torch").decompress(bad)
    hdr_len = 32 + 1 + 4  # 1-D shape, 4-byte dim
    bad = s.clone()
    bad[hdr_len + 2] = 7  # type byte of (group 1, chunk 0) out of range
    _expect_corrupt(bad)
    bad = s.clone()
    bad[-1] = 0  # last stream loses its end mark
    _expect_corrupt(bad)
    bad = s.clone()
    K = 2
    bad[hdr_len + 2 * K + 8 * K + 3] ^= 0x40  # cumulative size of (group 1, chunk 0) scrambled
    _expect_corrupt(bad)
    bad = s.clone()
    bad[hdr_len + 2 * K + 16 * K + 200000 + 3] ^= 0xFF  # inside the table description of the first coded block
    try:
        ZipNN(input_format="torch").decompress(bad)
    except RuntimeError as e:
        assert "corrupt" in str(e)
    torch.cuda.synchronize()
    _expect_corrupt(s[: s.numel() - 1000].clone())  # truncated
    # random bit flips inside the Huffman payload: either a clean error or wrong data, never a crash
    rng = np.random.default_rng(1)
    for _ in range(20):
        bad = s.clone()
        pos = int(rng.integers(s.numel() // 2 + 64, s.numel()))
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            ZipNN(input_format="torch").decompress(bad)
        except RuntimeError as e:
            assert "corrupt" in str(e)
        torch.cuda.synchronize()
    good = ZipNN(input_format="torch").decompress(s)
    assert torch.equal(good.view(torch.int16), t.view(torch.int16))


def test_reference_stress_sizes_round_trip():
    """The reference's own stress sizes (tests/simple_stress_tests.py:19-70): bf16 rand*2-1 and
    random bytes at 255/256/257/511/512/513/1024 KiB and about 1 and 2 MiB."""
    g = torch.Generator().manual_seed(0)
    for kib in (255, 256, 257, 511, 512, 513, 1024, 1025, 2049):
        n = kib * 1024 // 2
        t = (torch.rand(n, generator=g) * 2 - 1).to(torch.bfloat16)
        s = ZipNN(input_format="torch").compress(t.cuda())
        plan = ZipNN(input_format="torch").plan(t)
        want = O.zipnn_compress(plan["header"], np.frombuffer(raw_bytes(t), dtype=np.uint8), 2, 1, 10, 262144, 0.95, threads=4)
        assert s.cpu().numpy().tobytes() == want.tobytes()
        assert torch.equal(ZipNN(input_format="torch").decompress(s).cpu().view(torch.int16), t.view(torch.int16))
        b = torch.randint(0, 256, (kib * 1024,), generator=g, dtype=torch.uint8).numpy().tobytes()
        zb = ZipNN(input_format="byte", bytearray_dtype="bfloat16")
        sb = zb.compress(b)
        assert bytes(ZipNN(input_format="byte", bytearray_dtype="bfloat16").decompress(bytes(sb))) == b


def test_streaming_and_delta_frames():
    g = torch.Generator().manual_seed(2)
    a = torch.randint(0, 256, (3 * 1024 * 1024 + 17,), generator=g, dtype=torch.uint8).numpy().tobytes()
    base = (torch.randn(len(a) // 2 + 1, generator=g) * 0.02).to(torch.bfloat16).view(torch.uint8).numpy().tobytes()[: len(a)]
    for sc in (1 << 19, 1 << 20):
        z = ZipNN(input_format="byte", bytearray_dtype="bfloat16", is_streaming=True, streaming_chunk=sc)
        s = z.compress(base)
        assert bytes(ZipNN(input_format="byte", bytearray_dtype="bfloat16", is_streaming=True, streaming_chunk=sc).decompress(bytes(s))) == base
        # each frame is an independent reference stream
        plan = ZipNN(input_format="byte", bytearray_dtype="bfloat16", is_streaming=True, streaming_chunk=sc).plan(base[:sc])
        first = O.zipnn_compress(plan["header"], np.frombuffer(base[:sc], dtype=np.uint8), 2, 1, 10, 262144, 0.95)
        assert bytes(s[: first.size]) == first.tobytes()
    zd = ZipNN(input_format="byte", bytearray_dtype="bfloat16", delta_compressed_type="byte")
    s = zd.compress(base, delta_second_data=a)
    out = ZipNN(input_format="byte", bytearray_dtype="bfloat16", delta_compressed_type="byte").decompress(bytes(s), delta_second_data=a)
    assert bytes(out) == base
    # the delta stream is the reference's: XOR on the host, then the plain codec (zipnn/zipnn.py:636-640)
    x = np.bitwise_xor(np.frombuffer(base, dtype=np.uint8), np.frombuffer(a, dtype=np.uint8))
    plan = ZipNN(input_format="byte", bytearray_dtype="bfloat16", delta_compressed_type="byte").plan(x.tobytes())
    want = O.zipnn_compress(plan["header"], x, 2, 1, 10, 262144, 0.95, threads=4)
    assert bytes(s) == want.tobytes()
    # device-resident operands: the XOR runs on the GPU and CUDA tensors come back
    d_base = torch.frombuffer(bytearray(base), dtype=torch.uint8).cuda()
    d_a = torch.frombuffer(bytearray(a), dtype=torch.uint8).cuda()
    sd = ZipNN(input_format="byte", bytearray_dtype="bfloat16", delta_compressed_type="byte").compress(d_base, delta_second_data=d_a)
    assert sd.is_cuda and sd.cpu().numpy().tobytes() == want.tobytes()
    od = ZipNN(input_format="byte", bytearray_dtype="bfloat16", delta_compressed_type="byte").decompress(sd, delta_second_data=d_a)
    assert od.is_cuda and od.cpu().numpy().tobytes() == base


# ------------------------------------------------------------------ size-independent properties at scale
def test_large_round_trip_properties():
    """256 MiB bf16: round trip is exact, the stream is deterministic, the chunk table is
    consistent with the stream length, and the ratio matches the reference's (0.6623)."""
    n = 128 * 1024 * 1024
    g = torch.Generator(device="cuda").manual_seed(1234)
    t = (torch.randn(n, generator=g, device="cuda") * 0.02).to(torch.bfloat16)
    z = ZipNN(input_format="torch")
    s1 = z.compress(t)
    s2 = ZipNN(input_format="torch").compress(t)
    assert torch.equal(s1, s2)
    ratio = s1.numel() / (2 * n)
    assert 0.655 < ratio < 0.670, ratio
    head = s1[:64].cpu().numpy()
    assert int.from_bytes(head[24:32].tobytes(), "little") == s1.numel()
    K = (2 * n + 262143) // 262144
    hdr_len = 32 + 1 + 5
    cum = s1[hdr_len + 2 * K: hdr_len + 2 * K + 16 * K].cpu().numpy().view(np.uint64).reshape(2, K)
    assert hdr_len + 18 * K + int(cum[0, -1]) + int(cum[1, -1]) == s1.numel()
    assert np.all(np.diff(cum[0].astype(np.int64)) == 131072)  # sign|mantissa plane is stored raw
    back = ZipNN(input_format="torch").decompress(s1)
    assert torch.equal(back.view(torch.int16), t.view(torch.int16))
    # spot-check one window of chunks against the oracle
    c0 = 301
    part = t[c0 * 131072: (c0 + 3) * 131072].cpu()
    plan = ZipNN(input_format="torch").plan(part)
    want = O.zipnn_compress(plan["header"], np.frombuffer(raw_bytes(part), dtype=np.uint8), 2, 1, 10, 262144, 0.95, threads=3)
    wcum = want[len(plan["header"]) + 6: len(plan["header"]) + 6 + 48].view(np.uint64).reshape(2, 3)
    assert int(wcum[1, -1]) == int(cum[1, c0 + 2] - cum[1, c0 - 1])


def test_pipelined_host_decompress(monkeypatch):
    """Host streams above a size threshold are decoded slab by slab on two CUDA streams, inside
    zipnn_b200_decompress_host (the environment knob shrinks the slabs so small inputs take that path)."""
    monkeypatch.setenv("ZIPNN_B200_HOST_SLAB_BYTES", str(3 * 262144))
    g = torch.Generator().manual_seed(5)
    for dt, n in ((torch.bfloat16, 5 * 131072 * 2 + 12345), (torch.float32, 11 * 65536 + 7), (torch.float8_e4m3fn, 9 * 131072 + 1)):
        t = (torch.randn(n, generator=g) * 0.02).to(dt)
        t[1000:200000] = 0  # RLE planes in some chunks
        s = ZipNN(input_format="torch").compress(t)
        back = ZipNN(input_format="torch").decompress(bytes(s))
        assert back.dtype == dt and raw_bytes(back) == raw_bytes(t)
        bad = bytearray(bytes(s))
        hdr_len = 32 + 1 + 1 + 4   # header + 1-D shape with a 4-byte dimension
        bad[hdr_len + 1] = 9        # a type byte out of range is always detectable
        with pytest.raises(RuntimeError, match="corrupt"):
            ZipNN(input_format="torch").decompress(bytes(bad))


@pytest.mark.gpu
def test_pipelined_host_compress(monkeypatch):
    """Large host inputs are compressed slab by slab inside zipnn_b200_compress_host (H2D of the next slab, D2H
    of group 0 of the previous one in flight together); the stream must equal the one-shot device stream."""
    g = torch.Generator().manual_seed(7)
    cases = ((torch.bfloat16, 5 * 131072 * 2 + 12345), (torch.float32, 11 * 65536 + 7), (torch.float16, 7 * 131072),
             (torch.float8_e4m3fn, 9 * 131072 + 1))
    for dt, n in cases:
        t = (torch.randn(n, generator=g) * 0.02).to(dt)
        t[1000:200000] = 0  # RLE planes in some chunks
        one_shot = ZipNN(input_format="torch").compress(t.cuda()).cpu().numpy().tobytes()
        with monkeypatch.context() as m:
            m.setenv("ZIPNN_B200_HOST_SLAB_BYTES", str(3 * 262144))
            piped = bytes(ZipNN(input_format="torch").compress(t))
            out = torch.empty(len(one_shot) + 4096, dtype=torch.uint8, pin_memory=True)
            piped_out = bytes(ZipNN(input_format="torch").compress(t.pin_memory(), out=out))
        assert piped == one_shot and piped_out == one_shot, str(dt)
        back = ZipNN(input_format="torch").decompress(piped)
        assert raw_bytes(back) == raw_bytes(t)


@pytest.mark.gpu
@pytest.mark.parametrize("slab", [None, 3 * 262144])
def test_host_exports_with_plain_malloc_buffers(slab, monkeypatch):
    """zipnn_b200_compress_host / zipnn_b200_decompress_host called the way a C extension standing in for the
    reference's zipnn_core module would (csrc/zipnn_core.c:401-417, 881-892): pageable malloc'ed host buffers,
    no torch in between.  Stream == oracle stream; round trip exact; also a mixed case where a later group is
    coded although the first is raw (the early-copy bet of the slab path is lost and repaired)."""
    if slab:
        monkeypatch.setenv("ZIPNN_B200_HOST_SLAB_BYTES", str(slab))
    L = _native.lib()
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(31)
    cases = []
    x = (rng.standard_normal(131072 * 7 + 333) * 0.02).astype(np.float32)
    cases.append((2, 1, 10, 262144, np.ascontiguousarray((x.view(np.uint32) >> 16).astype(np.uint16).view(np.uint8))))    # bf16
    cases.append((4, 1, 220, 262144, np.ascontiguousarray(x.view(np.uint8)[: 65536 * 4 * 5 + 8])))                          # fp32
    cases.append((1, 0, 10, 131072, rng.choice(6, 131072 * 5 + 1, p=[.5, .25, .12, .06, .04, .03]).astype(np.uint8)))       # fp8-like, all coded
    mixed = np.ascontiguousarray((x.view(np.uint32) >> 16).astype(np.uint16).view(np.uint8)).copy()
    mixed[0: 262144 * 2: 2] = 7           # the low byte plane of two chunks becomes compressible: group 0 is no longer all raw
    cases.append((2, 1, 10, 262144, mixed))
    for G, bits, bm, chunk, data in cases:
        n = data.size
        hdr = bytearray(32)
        hdr[0:2] = b"ZN"
        want = O.zipnn_compress(hdr, data, G, bits, bm, chunk, 0.95, threads=4)
        bound = _native.compress_bound(n, G, chunk, 32)
        p_in, p_out, p_back = libc.malloc(n), libc.malloc(bound), libc.malloc(n)
        try:
            C.memmove(p_in, data.ctypes.data, n)
            out_len = C.c_size_t(0)
            hbuf = (C.c_char * 32).from_buffer_copy(bytes(hdr))
            rc = L.zipnn_b200_compress_host(p_in, n, hbuf, 32, G, bits, bm, chunk, 0.95, p_out, bound, C.byref(out_len))
            assert rc == 0
            got = np.frombuffer((C.c_char * out_len.value).from_address(p_out), dtype=np.uint8)
            assert out_len.value == want.size and np.array_equal(got, want), (G, bits, n)
            rc = L.zipnn_b200_decompress_host(p_out + 32, out_len.value - 32, G, bits, bm, chunk, n, p_back)
            assert rc == 0
            back = np.frombuffer((C.c_char * n).from_address(p_back), dtype=np.uint8)
            assert np.array_equal(back, data)
            # a corrupt type byte is reported, not decoded
            bad = np.array(got, copy=True)
            bad[32 + 1 if G * ((n + chunk - 1) // chunk) > 1 else 32] = 5
            C.memmove(p_out, bad.ctypes.data, bad.size)
            assert L.zipnn_b200_decompress_host(p_out + 32, out_len.value - 32, G, bits, bm, chunk, n, p_back) == _native.E_CORRUPT
        finally:
            for q in (p_in, p_out, p_back):
                libc.free(q)


# ------------------------------------------------------------------ round-2 paths
def _two_coded_groups_fp32(nchunks, seed=3):
    """fp32 values that are exact bf16 numbers: byte groups 0 and 1 are all zero (RLE), groups 2 and 3
    are both Huffman-coded -> every chunk needs the general path (plane scratch)."""
    g = torch.Generator().manual_seed(seed)
    n = nchunks * 65536
    return (torch.randn(n, generator=g) * 0.02).to(torch.bfloat16).to(torch.float32)


@pytest.mark.gpu
@pytest.mark.parametrize("nchunks", [40, 150, 4000])
def test_general_chunks_beyond_the_slot_pool_are_decoded_in_stream(nchunks, monkeypatch):
    """More general-mode chunks than the default workspace has pool slots (64): the overflow kernel must
    decode the rest in stream order -- no E_CAPACITY round trip, nothing left unwritten.
    4000 chunks is past the sync-kernel threshold, so the one-thread-per-bitstream kernels run too."""
    x = _two_coded_groups_fp32(nchunks)
    raw = x.view(torch.uint8).numpy()
    plan = ZipNN(input_format="torch").plan(x)
    want = O.zipnn_compress(plan["header"], raw, 4, 1, 220, 262144, 0.95, threads=8)
    stream = ZipNN(input_format="torch").compress(x.cuda())
    assert np.array_equal(stream.cpu().numpy(), want)
    # C ABI directly, DEFAULT workspace, one call
    L = _native.lib()
    body = stream[len(plan["header"]):].contiguous()
    pad = torch.zeros(64 + body.numel() + 16, dtype=torch.uint8, device="cuda")
    pad[64: 64 + body.numel()] = body
    out = torch.full((raw.size,), 0xA5, dtype=torch.uint8, device="cuda")
    ws = torch.empty(_native.decompress_workspace_size(raw.size, 4, 262144), dtype=torch.uint8, device="cuda")
    rc = L.zipnn_b200_decompress(pad[64:].data_ptr(), body.numel(), 4, 1, 220, 262144, raw.size, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                 torch.cuda.current_stream().cuda_stream, 1)
    assert rc == 0
    assert np.array_equal(out.cpu().numpy(), raw)


@pytest.mark.gpu
def test_batch_decompress_matches_single_calls():
    """zipnn_b200_decompress_batch: tensors of mixed dtype and size (empty, tiny, ragged, multi-coded,
    one above the sync threshold) in one call == one call per tensor == the original bytes."""
    rng = np.random.default_rng(9)
    L = _native.lib()
    specs = [(torch.bfloat16, 0), (torch.bfloat16, 5), (torch.bfloat16, 300001), (torch.float32, 70001), (torch.float16, 131072 * 3),
             (torch.float8_e4m3fn, 200000), (torch.bfloat16, 131072 * 40), (torch.float32, 65536 * 3300)]
    tensors = [torch.from_numpy(rng.standard_normal(n, dtype=np.float32) * np.float32(0.02 if dt != torch.float8_e4m3fn else 0.5)).to(dt)
               for dt, n in specs]
    tensors.append(_two_coded_groups_fp32(120))
    streams, plans = [], []
    for t in tensors:
        z = ZipNN(input_format="torch")
        streams.append(z.compress(t.cuda()))
        plans.append(z._last_plan)
    bodies, outs = [], []
    arr = (_native.BatchItem * len(tensors))()
    for i, (t, s, p) in enumerate(zip(tensors, streams, plans)):
        hl = len(p["header"])
        buf = torch.zeros(64 + s.numel() - hl + 16, dtype=torch.uint8, device="cuda")
        buf[64: 64 + s.numel() - hl] = s[hl:]
        bodies.append(buf)
        n = t.numel() * t.element_size()
        out = torch.full((max(n, 1),), 0x5A, dtype=torch.uint8, device="cuda")
        outs.append(out)
        arr[i].d_body = buf.data_ptr() + 64
        arr[i].body_len = s.numel() - hl
        arr[i].num_buf, arr[i].bits_mode, arr[i].bytes_mode = p["num_buf"], p["bit_reorder"], p["byte_reorder"]
        arr[i].chunk, arr[i].orig = p["chunk"], n
        arr[i].d_out = out.data_ptr()
    wsz = C.c_size_t(0)
    assert L.zipnn_b200_decompress_batch_workspace_size(arr, len(tensors), C.byref(wsz)) == 0
    ws = torch.empty(wsz.value, dtype=torch.uint8, device="cuda")
    before = _native.launch_count()
    rc = L.zipnn_b200_decompress_batch(arr, len(tensors), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream, 1)
    assert rc == 0
    launched = _native.launch_count() - before
    assert launched <= 12, f"{launched} launches for {len(tensors)} tensors: the small ones must share launches"
    for t, out in zip(tensors, outs):
        n = t.numel() * t.element_size()
        assert out[:n].cpu().numpy().tobytes() == raw_bytes(t)
    # a corrupt member fails the batch
    bad = bodies[2].clone()
    bad[64 + 2] = 9
    arr[2].d_body = bad.data_ptr() + 64
    rc = L.zipnn_b200_decompress_batch(arr, len(tensors), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream, 1)
    assert rc == _native.E_CORRUPT
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_crafted_size_table_cannot_wrap():
    """A cumulative size near 2^64 in group 0 would wrap the base of group 1 and point its items in front of
    the body: the whole stream must be rejected, nothing dereferenced."""
    t = (torch.randn(131072 * 2) * 0.02).to(torch.bfloat16).cuda()    # 2 chunks
    s = ZipNN(input_format="torch").compress(t).clone()
    hdr_len, G, K = 32 + 1 + 4, 2, 2
    cum0_last = hdr_len + G * K + 8 * (0 * K + (K - 1))
    for evil in ((1 << 64) - 4096, (1 << 64) - 1, 1 << 63, (1 << 40)):
        bad = s.clone()
        bad[cum0_last: cum0_last + 8] = torch.frombuffer(bytearray(int(evil).to_bytes(8, "little")), dtype=torch.uint8).cuda()
        _expect_corrupt(bad)
        torch.cuda.synchronize()
    good = ZipNN(input_format="torch").decompress(s)
    assert torch.equal(good.view(torch.int16), t.view(torch.int16))


@pytest.mark.gpu
@pytest.mark.parametrize("sync_max", ["0", "1000000"])
def test_both_decoder_families_agree(sync_max, monkeypatch):
    """The per-bitstream CTAs (decode_sync.cuh) and the one-thread-per-bitstream kernels must give the same
    bytes on the same streams, whatever the size threshold says (ZIPNN_B200_SYNC_MAX forces either)."""
    monkeypatch.setenv("ZIPNN_B200_SYNC_MAX", sync_max)
    rng = np.random.default_rng(21)
    for dt, n in ((torch.bfloat16, 131072 * 9 + 77), (torch.float32, 65536 * 5 + 3), (torch.float16, 131072 * 4), (torch.float8_e4m3fn, 131072 * 3 + 1),
                  (torch.bfloat16, 2048), (torch.bfloat16, 131072 * 70)):
        x = torch.from_numpy(rng.standard_normal(n, dtype=np.float32) * np.float32(0.02 if dt != torch.float8_e4m3fn else 0.5)).to(dt)
        s = ZipNN(input_format="torch").compress(x.cuda())
        back = ZipNN(input_format="torch").decompress(s)
        assert torch.equal(back.cpu().view(torch.uint8), x.view(torch.uint8))
    y = _two_coded_groups_fp32(70)
    s = ZipNN(input_format="torch").compress(y.cuda())
    assert torch.equal(ZipNN(input_format="torch").decompress(s).cpu().view(torch.uint8), y.view(torch.uint8))
    # small chunks (many bitstreams of a few hundred symbols)
    z = ZipNN(input_format="torch", compression_chunk=4096)
    x = (torch.randn(50001) * 0.02).to(torch.bfloat16)
    s = z.compress(x.cuda())
    assert torch.equal(ZipNN(input_format="torch").decompress(s).cpu().view(torch.uint8), x.view(torch.uint8))
