"""Multi-GPU host logic on CPU: world_size 2 over gloo.  The codec calls are replaced by the
oracle so that only the sharding/assembly code of zipnn_b200.sharded is under test: the merged
stream must be byte-identical to the single-process reference stream, and the scatter must hand
every rank a self-contained stream for its chunk range."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from golden_inputs import raw_bytes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_codecs():
    from oracle import oracle as O

    def comp(flat, hdr, G, bits, bm, chunk, thr):
        return torch.from_numpy(O.zipnn_compress(hdr, flat.numpy(), G, bits, bm, chunk, thr, threads=2).copy())

    def dec(body, G, bits, bm, chunk, orig):
        return torch.from_numpy(O.zipnn_decompress(body.numpy(), G, bits, bm, chunk, orig, threads=2).copy())
    return comp, dec


def _worker(rank, world, port, dtype_name, n_elems, chunk, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from zipnn_b200 import ZipNN
        from zipnn_b200.sharded import ShardedZipNN, byte_range
        dtype = getattr(torch, dtype_name)
        g = torch.Generator().manual_seed(77)
        full = (torch.randn(n_elems, generator=g) * 0.02).to(dtype)
        if n_elems > 1000:
            full[100:400] = 0  # a run of equal exponents/mantissas inside one chunk
        esz = full.element_size()
        nbytes = n_elems * esz
        eff_chunk = min(chunk, 131072) if esz == 1 else chunk
        b0, b1 = byte_range(nbytes, eff_chunk, rank, world)
        local = full[b0 // esz: b1 // esz].clone()
        comp, dec = _oracle_codecs()
        z = ShardedZipNN(compress_local=comp, decompress_local=dec, compression_chunk=chunk)
        stream = z.compress(local, global_shape=tuple(full.shape), dst=0)
        if rank == 0:
            plan = ZipNN(input_format="torch", compression_chunk=chunk).plan(full)
            want = O.zipnn_compress(plan["header"], np.frombuffer(raw_bytes(full), dtype=np.uint8), plan["num_buf"],
                                    plan["bit_reorder"], plan["byte_reorder"], plan["chunk"], plan["threshold"], threads=2)
            assert stream is not None and stream.numel() == want.size, (stream.numel(), want.size)
            assert np.array_equal(stream.numpy(), want)
        else:
            assert stream is None
        back = z.decompress(stream if rank == 0 else None, src=0)
        assert back.dtype == dtype
        assert raw_bytes(back) == raw_bytes(local), f"rank {rank} shard mismatch"
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name,n_elems,chunk", [
    ("bfloat16", 5 * 131072 + 777, 262144),   # 6 chunks, ragged tail, odd split 3/3
    ("float32", 3 * 65536 + 5, 262144),       # 4 chunks
    ("float16", 40000, 4096),                 # many small chunks
    ("float8_e4m3fn", 200001, 262144),        # one group, 128 KiB chunks
    ("bfloat16", 1000, 262144),               # a single chunk: rank 1 owns nothing
])
def test_gather_and_scatter_over_gloo(dtype_name, n_elems, chunk):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dtype_name, n_elems, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_chunk_ranges_cover_everything():
    from zipnn_b200.sharded import byte_range, chunk_range
    for K in (0, 1, 2, 7, 8, 9, 65536):
        for world in (1, 2, 4, 8):
            spans = [chunk_range(K, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == K
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    assert byte_range(1000, 256, 1, 2) == (512, 1000)


def _misaligned_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from zipnn_b200.sharded import ShardedZipNN
        comp, dec = _oracle_codecs()
        g = torch.Generator().manual_seed(3)
        # an "even split" that is not chunk aligned: rank 0 holds 1.5 chunks
        local = (torch.randn(196608 if rank == 0 else 65536, generator=g) * 0.02).to(torch.bfloat16)
        z = ShardedZipNN(compress_local=comp, decompress_local=dec)
        try:
            z.compress(local, global_shape=(196608 + 65536,), dst=0)
            q.put((rank, "no error"))
        except ValueError as e:
            q.put((rank, "ok" if "whole number" in str(e) else "wrong message " + str(e)))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_misaligned_shards_are_refused():
    """a shard that is not a whole number of chunks would yield a stream with a short chunk in
    the middle, silently different from the single-GPU stream."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_misaligned_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
