#!/usr/bin/env python3
"""Synthetic sweep (BASELINE.json configs[4]): dtype x size, device-resident compress and
decompress GB/s (uncompressed bytes / time, CUDA events, median of `reps`), ratio, and a
byte-exact round-trip check per point.
usage: python tools/sweep.py [--max-gib 4] [--dtypes bfloat16,float16,float32,float8_e4m3fn]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_tensor  # noqa: E402
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-gib", type=float, default=4.0)
    ap.add_argument("--dtypes", default="bfloat16,float16,float32,float8_e4m3fn")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    sizes = [s for s in (1 << 20, 16 << 20, 256 << 20, 1 << 30, 4 << 30, 16 << 30) if s <= args.max_gib * (1 << 30)]
    rows = []
    for dn in args.dtypes.split(","):
        dt = getattr(torch, dn)
        for n in sizes:
            t = make_tensor(n, dt, torch.device("cuda", 0), 1234)
            for _ in range(2):
                s = ZipNN(input_format="torch").compress(t)
                d = ZipNN(input_format="torch").decompress(s)
            assert torch.equal(d.view(torch.uint8), t.view(torch.uint8))
            del d
            tc, s = timed(lambda: ZipNN(input_format="torch").compress(t), args.reps)
            td, d = timed(lambda: ZipNN(input_format="torch").decompress(s), args.reps)
            row = dict(dtype=dn, bytes=n, ratio=round(s.numel() / n, 4), compress_gbs=round(n / tc / 1e6, 1),
                       decompress_gbs=round(n / td / 1e6, 1), compress_ms=round(tc, 3), decompress_ms=round(td, 3))
            rows.append(row)
            print(json.dumps(row), flush=True)
            del t, s, d
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
