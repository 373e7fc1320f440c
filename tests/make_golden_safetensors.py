#!/usr/bin/env python3
"""Write tests/golden/ref_model.znn.safetensors with the UNMODIFIED reference script
(scripts/zipnn_compress_safetensors.py), from a small synthetic checkpoint whose tensors are
the reference's own test pattern (tests/simple_stress_tests.py:215-255: half constant 42.0,
half randn) plus an int tensor that must stay untouched.  Build container only."""
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, "/root/reference/scripts")
sys.path.insert(0, HERE)

import torch  # noqa: E402
from safetensors.torch import save_file  # noqa: E402

from golden_safetensors_inputs import make_checkpoint  # noqa: E402
import zipnn_compress_safetensors as ref_script  # noqa: E402


def main():
    tensors = make_checkpoint()
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "ref_model.safetensors")
        save_file(tensors, src, {"format": "pt"})
        ref_script.compress_safetensors_file(src, force=True, threads=4)
        out = os.path.join(d, "ref_model.znn.safetensors")
        data = open(out, "rb").read()
    with open(os.path.join(HERE, "golden", "ref_model.znn.safetensors"), "wb") as f:
        f.write(data)
    print("wrote", len(data), "bytes")


if __name__ == "__main__":
    main()
