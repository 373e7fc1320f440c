#!/usr/bin/env python3
"""Per-kernel device times of one compress + decompress at a given size (library timing hooks).
usage: python tools/kernel_times.py [size_mib] [dtype] [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_tensor  # noqa: E402
from zipnn_b200 import ZipNN, _native  # noqa: E402

mib = float(sys.argv[1]) if len(sys.argv) > 1 else 1024.0
dtype = getattr(torch, sys.argv[2]) if len(sys.argv) > 2 else torch.bfloat16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
t = make_tensor(int(mib * (1 << 20)), dtype, "cuda", 1234)
for _ in range(2):
    s = ZipNN(input_format="torch").compress(t)
    d = ZipNN(input_format="torch").decompress(s)
assert torch.equal(d.view(torch.uint8), t.view(torch.uint8))
_native.timing_enable(True)
for _ in range(reps):
    s = ZipNN(input_format="torch").compress(t)
    d = ZipNN(input_format="torch").decompress(s)
kt = _native.timing_collect()
_native.timing_enable(False)
print(json.dumps({"mib": mib, "dtype": str(dtype), "kernels_ms": {k: round(ms / max(c, 1), 4) for k, (ms, c) in kt.items() if c}}))
