#!/usr/bin/env python3
"""Generate tests/golden/ from the UNMODIFIED Python reference.

Runs only in the build container: it imports `zipnn` from /root/reference with
the reference's own C extension compiled into oracle/_ref (make -C oracle ref).
The fixtures it writes are what travels to the GPU box:

  tests/golden/manifest.json   one record per case: how to regenerate the input
                               (seeded numpy), input sha256, stream length + sha256,
                               the ZipNN ctor kwargs, and -- for small cases -- the
                               file holding the full reference stream.
  tests/golden/<case>.znn      the reference's compressed stream, byte for byte.

The reference tests hold no golden vectors, so these are the
known-answer tests for the path (SURVEY.md section 8c lists the same cases).

usage:  python tests/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from zipnn import ZipNN  # noqa: E402  (the reference)

from golden_inputs import make_input  # noqa: E402

OUT = os.path.join(HERE, "golden")
FULL_STREAM_LIMIT = 96 * 1024


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()


CASES = [
    # name, input spec, ctor kwargs
    # ---- torch format
    ("bf16_zeros_ones", dict(gen="zeros_ones", dtype="bfloat16", n=1048576), dict(input_format="torch")),
    ("bf16_1m", dict(gen="randn", dtype="bfloat16", n=1048576, sigma=0.02), dict(input_format="torch")),
    ("bf16_ragged", dict(gen="randn", dtype="bfloat16", n=1060921, sigma=0.02), dict(input_format="torch")),
    ("fp16_1m", dict(gen="randn", dtype="float16", n=1048576, sigma=0.02), dict(input_format="torch")),
    ("fp32_512k", dict(gen="randn", dtype="float32", n=524288, sigma=0.02), dict(input_format="torch")),
    ("fp32_ragged", dict(gen="randn", dtype="float32", n=525065, sigma=0.02), dict(input_format="torch")),
    ("fp8e4m3_1m", dict(gen="randn", dtype="float8_e4m3fn", n=1048576, sigma=0.5), dict(input_format="torch")),
    ("fp8e5m2_300k", dict(gen="randn", dtype="float8_e5m2", n=300001, sigma=0.5), dict(input_format="torch")),
    ("fp32_from_bf16", dict(gen="randn_bf16_as_fp32", dtype="float32", n=300000, sigma=0.02), dict(input_format="torch")),
    ("bf16_2d_shape", dict(gen="randn", dtype="bfloat16", n=70000 * 3, sigma=0.02, shape=[70000, 3]), dict(input_format="torch")),
    ("bf16_uniform", dict(gen="rand_pm1", dtype="bfloat16", n=262144 + 77, sigma=1.0), dict(input_format="torch")),
    # ---- the 128 MiB known answer (512 chunks; stream sha 1e8d0f8b62398fb7)
    ("bf16_128mib", dict(gen="randn", dtype="bfloat16", n=67108864, sigma=0.02), dict(input_format="torch")),
    # ---- small ones whose full stream is committed
    ("bf16_small", dict(gen="randn", dtype="bfloat16", n=40000, sigma=0.02), dict(input_format="torch")),
    ("fp16_small", dict(gen="randn", dtype="float16", n=40000, sigma=0.02), dict(input_format="torch")),
    ("fp32_small", dict(gen="randn", dtype="float32", n=20000, sigma=0.02), dict(input_format="torch")),
    ("fp8_small", dict(gen="randn", dtype="float8_e4m3fn", n=50000, sigma=0.5), dict(input_format="torch")),
    ("bf16_chunk4k", dict(gen="randn", dtype="bfloat16", n=30001, sigma=0.02),
     dict(input_format="torch", compression_chunk=4096)),
    ("fp32_chunk64k", dict(gen="randn", dtype="float32", n=50001, sigma=0.02),
     dict(input_format="torch", compression_chunk=65536)),
    # ---- byte format (SURVEY 8c tables 1 + 2)
    ("byte_bf16_huffman", dict(gen="randn", dtype="bfloat16", n=1048576, sigma=0.02),
     dict(input_format="byte", bytearray_dtype="bfloat16", method="HUFFMAN")),
    ("byte_fp8_nibble", dict(gen="choice", dtype="uint8", n=1 << 20, p=[.5, .25, .12, .06, .04, .03]),
     dict(input_format="byte", bytearray_dtype="float8_e4m3fn", method="HUFFMAN")),
    ("byte_fp8_2sym", dict(gen="choice", dtype="uint8", n=1 << 20, p=[.9, .1]),
     dict(input_format="byte", bytearray_dtype="float8_e4m3fn", method="HUFFMAN")),
    ("byte_fp8_allraw", dict(gen="bytes", dtype="uint8", n=1 << 20),
     dict(input_format="byte", bytearray_dtype="float8_e4m3fn", method="HUFFMAN")),
    ("byte_fp8_threshold", dict(gen="half_zero", dtype="uint8", n=1 << 20, frac=0.52),
     dict(input_format="byte", bytearray_dtype="float8_e4m3fn", method="HUFFMAN")),
    ("byte_bf16_tiny2", dict(gen="randn", dtype="bfloat16", n=1, sigma=0.02),
     dict(input_format="byte", bytearray_dtype="bfloat16", method="HUFFMAN")),
    ("byte_bf16_tiny4", dict(gen="randn", dtype="bfloat16", n=2, sigma=0.02),
     dict(input_format="byte", bytearray_dtype="bfloat16", method="HUFFMAN")),
    ("byte_bf16_tiny24", dict(gen="randn", dtype="bfloat16", n=12, sigma=0.02),
     dict(input_format="byte", bytearray_dtype="bfloat16", method="HUFFMAN")),
    ("byte_bf16_tiny26", dict(gen="randn", dtype="bfloat16", n=13, sigma=0.02),
     dict(input_format="byte", bytearray_dtype="bfloat16", method="HUFFMAN")),
    ("byte_bf16_4096", dict(gen="randn", dtype="bfloat16", n=2048, sigma=0.02),
     dict(input_format="byte", bytearray_dtype="bfloat16", method="HUFFMAN")),
    ("byte_fp32_small", dict(gen="randn", dtype="float32", n=9000, sigma=0.02),
     dict(input_format="byte", bytearray_dtype="float32", method="HUFFMAN")),
    ("byte_fp16_small", dict(gen="randn", dtype="float16", n=9001, sigma=0.02),
     dict(input_format="byte", bytearray_dtype="float16", method="AUTO")),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = []
    for name, spec, kw in CASES:
        data = make_input(spec)  # torch tensor (torch format) or bytes (byte format)
        if kw["input_format"] == "torch":
            raw = data.contiguous().view(torch.uint8).numpy().tobytes() if data.dtype != torch.uint8 else data.numpy().tobytes()
            arg = data.clone()  # the reference rotates its input in place
        else:
            raw = data if isinstance(data, (bytes, bytearray)) else data.contiguous().view(torch.uint8).numpy().tobytes()
            arg = bytearray(raw)
        z = ZipNN(**kw)
        stream = bytes(z.compress(arg))
        # self-check with the reference decoder
        back = ZipNN(**kw).decompress(stream)
        if kw["input_format"] == "torch":
            assert back.dtype == data.dtype and tuple(back.shape) == tuple(data.shape)
            assert back.contiguous().view(torch.uint8).numpy().tobytes() == raw, name
        else:
            assert bytes(back) == raw, name
        rec = dict(name=name, input=spec, ctor=kw, input_sha256=sha(raw), input_len=len(raw),
                   stream_len=len(stream), stream_sha256=sha(stream), header_hex=stream[:32].hex())
        if len(stream) <= FULL_STREAM_LIMIT:
            fn = name + ".znn"
            with open(os.path.join(OUT, fn), "wb") as f:
                f.write(stream)
            rec["stream_file"] = fn
        manifest.append(rec)
        print(f"{name:24s} in={len(raw):9d} out={len(stream):9d} ratio={len(stream)/max(len(raw),1):.4f} sha={rec['stream_sha256'][:16]}")
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(dict(reference="zipnn/zipnn v0.5.3 (0e9beed), built -O3 from /root/reference",
                       numpy=np.__version__, torch=torch.__version__, cases=manifest), f, indent=1)
    print("wrote", len(manifest), "cases")


if __name__ == "__main__":
    main()
