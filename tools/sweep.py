#!/usr/bin/env python3
"""Synthetic sweep (BASELINE.json configs[4]): dtype x size, device-resident compress and decompress
GB/s (uncompressed bytes / time, CUDA events, median of `reps`), stream ratio, achieved fraction of the
HBM roofline ((N + C) / t / peak), a byte-exact round-trip check per point and -- with
--cpu -- the reference's own C path on the host cores beside it (bounded sample of the same bytes).

  python tools/sweep.py [--max-gib 16] [--dtypes bfloat16,float16,float32,float8_e4m3fn] [--cpu]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py ...
      one shard of the given size per GPU (weak scaling); the line reports the whole-job rate (sum of
      bytes / slowest rank).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import cpu_reference_codec, make_tensor, peaks  # noqa: E402
from zipnn_b200 import ZipNN  # noqa: E402


def timed(fn, reps):
    ts = []
    out = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], out


def cpu_column(t, sample_bytes, threads):
    """Reference C path (oracle/_ref) or the oracle port on the first `sample_bytes` of the tensor."""
    layout = {torch.bfloat16: (2, 1, 10, 262144), torch.float16: (2, 0, 10, 262144), torch.float32: (4, 1, 220, 262144),
              torch.float8_e4m3fn: (1, 1, 10, 131072), torch.float8_e5m2: (1, 1, 10, 131072)}[t.dtype]
    kind, comp, dec, release = cpu_reference_codec(*layout)
    raw = bytearray(t.view(torch.uint8)[:sample_bytes].cpu().numpy().tobytes())
    work = bytearray(raw)
    t0 = time.perf_counter()
    s = comp(work, threads)
    t1 = time.perf_counter()
    d = dec(s, len(raw), threads)
    t2 = time.perf_counter()
    ok = bytes(d) == bytes(raw)
    if isinstance(d, memoryview):
        release(d)
    if isinstance(s, memoryview):
        release(s)
    n = len(raw)
    return {"kind": kind, "threads": threads, "sample_bytes": n, "compress_gbs": round(n / (t1 - t0) / 1e9, 3),
            "decompress_gbs": round(n / (t2 - t1) / 1e9, 3), "exact": ok}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-gib", type=float, default=4.0)
    ap.add_argument("--min-mib", type=float, default=1.0)
    ap.add_argument("--dtypes", default="bfloat16,float16,float32,float8_e4m3fn")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--cpu-sample-mib", type=float, default=1024.0)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peak, _ = peaks()
    sizes = [s for s in (1 << 20, 16 << 20, 256 << 20, 1 << 30, 4 << 30, 16 << 30) if args.min_mib * (1 << 20) <= s <= args.max_gib * (1 << 30)]
    for dn in args.dtypes.split(","):
        dt = getattr(torch, dn)
        for n in sizes:
            t = make_tensor(n, dt, dev, 1234 + rank)
            for _ in range(2):
                s = ZipNN(input_format="torch").compress(t)
                d = ZipNN(input_format="torch").decompress(s)
            exact = bool(torch.equal(d.view(torch.uint8), t.view(torch.uint8)))
            del d
            if world > 1:
                dist.barrier()
            tc, s = timed(lambda: ZipNN(input_format="torch").compress(t), args.reps)
            td, d = timed(lambda: ZipNN(input_format="torch").decompress(s), args.reps)
            C = s.numel()
            if world > 1:
                v = torch.tensor([tc, td], device=dev, dtype=torch.float64)
                dist.all_reduce(v, op=dist.ReduceOp.MAX)
                tc, td = [float(x) for x in v.tolist()]
            row = dict(dtype=dn, bytes_per_gpu=n, n_gpus=world, ratio=round(C / n, 4), exact=exact,
                       compress_gbs=round(world * n / tc / 1e6, 1), decompress_gbs=round(world * n / td / 1e6, 1),
                       compress_ms=round(tc, 3), decompress_ms=round(td, 3),
                       compress_roofline=round((n + C) / tc / 1e6 / peak, 4), decompress_roofline=round((n + C) / td / 1e6 / peak, 4))
            if args.cpu and rank == 0:
                cores = os.cpu_count() or 1
                row["cpu_reference"] = cpu_column(t, int(min(n, args.cpu_sample_mib * (1 << 20))), cores)
            if rank == 0:
                print(json.dumps(row), flush=True)
            del t, s, d
            torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
