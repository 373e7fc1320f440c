"""`zipnn_hf()` -- transformers plugin: checkpoints stored as `.znn` files (the reference's whole-file
streaming format, zipnn/zipnn.py:612-635) are decoded when a model is loaded.

Mirrors reference zipnn/zipnn.py:1221-1565 in what it does for the user:
  * `transformers.modeling_utils.load_state_dict` accepts `*.znn` files (`model.safetensors.znn`,
    `pytorch_model.bin.znn`, shards): the file is decoded -- on the GPU, all 1 MiB frames in one batched
    call (`ZipNN(is_streaming=True).decompress`) -- and the state dict is built from the decoded bytes; with
    `replace_local_file=True` the decoded file replaces the `.znn` one on disk (symlinked hub caches included)
    and the weight index json is rewritten, as the reference does (:1287-1313).
  * `PreTrainedModel.from_pretrained` on a local directory whose weights exist only as `.znn` files finds them.
    The reference re-implements the hub file resolution of transformers 4.x for that (:1409-1560), with private
    imports that transformers 5 no longer has (`is_torch_greater_or_equal`, ...); here the directory's `.znn`
    weight files are decoded to their plain names first (or to a temporary directory next to the config when
    `replace_local_file=False`), which is independent of transformers' internals.  Repositories that hold
    ONLY `.znn` files on the hub must be downloaded first (`huggingface_hub.snapshot_download`) -- the one
    difference from the reference's surface.
"""
from __future__ import annotations

import json
import os
import shutil
import tempfile
from io import BytesIO

import torch

from .zipnn import ZipNN

_WEIGHT_SUFFIXES = (".safetensors.znn", ".bin.znn", ".pt.znn", ".pth.znn", ".ckpt.znn")
_INDEX_NAMES = ("model.safetensors.index.json", "pytorch_model.bin.index.json")


def decompress_znn_file(path: str) -> bytes:
    """Decoded bytes of a `.znn` whole-file stream."""
    with open(path, "rb") as f:
        data = f.read()
    return bytes(ZipNN(is_streaming=True).decompress(data))


def _replace_in_file(file_path: str, old: str, new: str) -> None:
    with open(file_path, "r") as f:
        text = f.read()
    with open(file_path, "w") as f:
        f.write(text.replace(old, new))


def _materialise(checkpoint_file: str) -> str:
    """`x.znn` -> `x` on disk (reference :1287-1313): through the symlink when the file lives in a hub cache,
    and with the weight index pointing at the decoded name."""
    output_file = checkpoint_file[: -len(".znn")]
    snapshot = os.path.dirname(checkpoint_file)
    if not os.path.exists(output_file):
        data = decompress_znn_file(checkpoint_file)
        if os.path.islink(checkpoint_file):
            blob = os.path.join(snapshot, os.readlink(checkpoint_file))
            with open(output_file + ".tmp", "wb") as f:
                f.write(data)
            os.replace(output_file + ".tmp", blob)
            os.symlink(os.path.relpath(blob, snapshot), output_file)
        else:
            with open(output_file + ".tmp", "wb") as f:
                f.write(data)
            os.replace(output_file + ".tmp", output_file)
    os.remove(checkpoint_file)
    base = os.path.basename(output_file)
    for idx in _INDEX_NAMES:
        ip = os.path.join(snapshot, idx)
        if os.path.exists(ip):
            target = os.path.join(snapshot, os.readlink(ip)) if os.path.islink(ip) else ip
            _replace_in_file(target, base + ".znn", base)
    return output_file


def _state_dict_from_bytes(checkpoint_file: str, data: bytes, map_location, weights_only):
    if checkpoint_file.endswith(".safetensors.znn"):
        from safetensors.torch import load
        hlen = int.from_bytes(data[:8], "little")
        meta = json.loads(data[8: 8 + hlen]).get("__metadata__", {})
        if meta.get("format") not in ("pt", "tf", "flax", "mlx"):
            raise OSError(f"The safetensors archive passed at {checkpoint_file} does not contain the valid metadata. Make sure "
                          "you save your model with the `save_pretrained` method.")
        sd = load(data)
        if map_location not in (None, "cpu", torch.device("cpu")) and map_location != "meta":
            sd = {k: v.to(map_location) for k, v in sd.items()}
        return sd
    return torch.load(BytesIO(data), map_location=map_location or "cpu", weights_only=weights_only)


def zipnn_hf(replace_local_file: bool = False):
    """Patch transformers in this process (reference zipnn/zipnn.py:1221)."""
    try:
        import transformers
        from transformers import modeling_utils
        from transformers.modeling_utils import PreTrainedModel
    except ImportError as exc:
        raise ImportError("Hugging Face Transformers library is not installed. Please install it to use ZipNN compression.") from exc

    if getattr(modeling_utils, "_zipnn_b200_patched", False):
        return
    original_load_state_dict = modeling_utils.load_state_dict

    def custom_load_state_dict(checkpoint_file, *args, **kwargs):
        cf = os.fspath(checkpoint_file)
        if cf.endswith(".znn"):
            print(f"Decompressing {os.path.basename(cf)}")
            if replace_local_file:
                return original_load_state_dict(_materialise(cf), *args, **kwargs)
            map_location = kwargs.get("map_location", args[0] if args and not isinstance(args[0], bool) else "cpu")
            weights_only = kwargs.get("weights_only", True)
            return _state_dict_from_bytes(cf, decompress_znn_file(cf), map_location, weights_only)
        if not os.path.exists(cf) and os.path.exists(cf + ".znn"):
            return custom_load_state_dict(cf + ".znn", *args, **kwargs)
        return original_load_state_dict(checkpoint_file, *args, **kwargs)

    modeling_utils.load_state_dict = custom_load_state_dict

    original_from_pretrained = PreTrainedModel.from_pretrained.__func__

    def custom_from_pretrained(cls, pretrained_model_name_or_path, *model_args, **kwargs):
        path = os.fspath(pretrained_model_name_or_path) if isinstance(pretrained_model_name_or_path, (str, os.PathLike)) else None
        tmp = None
        if path and os.path.isdir(path):
            znn = [f for f in os.listdir(path) if f.endswith(_WEIGHT_SUFFIXES)]
            if znn and replace_local_file:
                for f in znn:
                    _materialise(os.path.join(path, f))
            elif znn:
                # decoded copies in a scratch directory that mirrors the checkpoint (everything else linked)
                tmp = tempfile.mkdtemp(prefix="zipnn_b200_hf_")
                for f in os.listdir(path):
                    src = os.path.join(path, f)
                    if f in znn:
                        with open(os.path.join(tmp, f[: -len(".znn")]), "wb") as out:
                            out.write(decompress_znn_file(src))
                    elif f in _INDEX_NAMES:
                        shutil.copy(src, os.path.join(tmp, f))
                        for z in znn:
                            _replace_in_file(os.path.join(tmp, f), z, z[: -len(".znn")])
                    else:
                        os.symlink(os.path.abspath(src), os.path.join(tmp, f))
                pretrained_model_name_or_path = tmp
        try:
            return original_from_pretrained(cls, pretrained_model_name_or_path, *model_args, **kwargs)
        finally:
            if tmp:
                shutil.rmtree(tmp, ignore_errors=True)

    PreTrainedModel.from_pretrained = classmethod(custom_from_pretrained)
    modeling_utils._zipnn_b200_patched = True
    _ = transformers
