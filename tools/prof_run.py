"""Small driver for ncu captures: one compress + one decompress of a bf16 tensor.
usage: python tools/prof_run.py [size_gib] [dtype]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_tensor  # noqa: E402
from zipnn_b200 import ZipNN  # noqa: E402

size = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
dtype = getattr(torch, sys.argv[2]) if len(sys.argv) > 2 else torch.bfloat16
t = make_tensor(int(size * (1 << 30)), dtype, torch.device("cuda", 0), 1234)
for _ in range(2):
    s = ZipNN(input_format="torch").compress(t)
    d = ZipNN(input_format="torch").decompress(s)
torch.cuda.synchronize()
assert torch.equal(d.view(torch.uint8), t.view(torch.uint8))
print("ok", s.numel() / (t.numel() * t.element_size()))
