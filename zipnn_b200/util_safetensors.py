"""safetensors container metadata for compressed tensors.

Mirrors reference zipnn/util_safetensors.py:9-58: a compressed tensor is stored
under its own name as uint8[stream_len]; the file-level metadata key
`znn_compressed_vectors` holds JSON {name: {"dtype": "...", "shape": "[...]"}}.
"""
import json
from typing import Dict

import torch

METADATA_KEY = "znn_compressed_vectors"
COMPRESSION_METHOD = "HUFFMAN"
COMPRESSED_DTYPE = torch.uint8


def build_compressed_tensor_info(uncompressed_tensor: torch.Tensor) -> Dict[str, str]:
    dtype = str(uncompressed_tensor.dtype)
    if dtype.startswith("torch."):
        dtype = dtype[len("torch."):]
    return {"dtype": dtype, "shape": str(list(uncompressed_tensor.shape))}


def set_compressed_tensors_metadata(compressed_tensor_infos, metadata):
    """Reference quirk kept: a file that had NO metadata dict gets none (util_safetensors.py:41-43),
    so callers that want the key must pass a dict (our writer always does)."""
    if metadata is not None:
        metadata[METADATA_KEY] = json.dumps(compressed_tensor_infos)


def get_compressed_tensors_metadata(metadata) -> Dict[str, Dict[str, str]]:
    if metadata:
        return json.loads(metadata.get(METADATA_KEY) or "{}")
    return {}
