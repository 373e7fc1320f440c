"""zipnn_b200 -- the ZipNN encode/decode hot path on B200 (sm_100a).

Drop-in for the reference package surface that sits on the hot path
(reference zipnn/__init__.py:1):  `ZipNN`, `zipnn_safetensors`.  `zipnn_hf` (the
transformers monkey-patch, zipnn/zipnn.py:1221-1565) is file plumbing outside the path and
is not provided.
"""
from .zipnn import DecodePipe, ZipNN
from .safetensors_io import (SafeOpen, compress_safetensors_file, decompress_safetensors_file,
                             decompress_safetensors_tensor, load_file, zipnn_safetensors)


def zipnn_hf(*args, **kwargs):
    raise NotImplementedError("zipnn_hf() is outside the B200 hot path (see DESIGN.md, 'out of scope'); "
                              "use zipnn_safetensors() for the vLLM / safetensors load path")


__all__ = ["ZipNN", "zipnn_safetensors", "SafeOpen", "compress_safetensors_file",
           "decompress_safetensors_file", "decompress_safetensors_tensor", "load_file", "DecodePipe", "zipnn_hf"]
