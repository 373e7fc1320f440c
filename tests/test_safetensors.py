"""The safetensors load path (SURVEY section 8 '* surface only', 8f N1): SafeOpen, the
zipnn_safetensors() patch, and the .znn.safetensors writer/reader.  Mirrors the reference's own
safetensors test (tests/simple_stress_tests.py:215-263) and adds interoperability with a file
written by the reference script (tests/golden/ref_model.znn.safetensors)."""
import json
import os

import pytest
import torch
from safetensors import safe_open
from safetensors.torch import save_file

from golden_inputs import raw_bytes
from golden_safetensors_inputs import make_checkpoint

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "ref_model.znn.safetensors")


def _same(a, b):
    return a.dtype == b.dtype and tuple(a.shape) == tuple(b.shape) and raw_bytes(a.cpu()) == raw_bytes(b.cpu())


def test_metadata_helpers_cpu():
    from zipnn_b200.util_safetensors import (build_compressed_tensor_info, get_compressed_tensors_metadata,
                                             set_compressed_tensors_metadata)
    t = torch.zeros(3, 5, dtype=torch.bfloat16)
    info = build_compressed_tensor_info(t)
    assert info == {"dtype": "bfloat16", "shape": "[3, 5]"}
    meta = {"format": "pt"}
    set_compressed_tensors_metadata({"w": info}, meta)
    assert json.loads(meta["znn_compressed_vectors"]) == {"w": info}
    assert get_compressed_tensors_metadata(meta) == {"w": info}
    assert get_compressed_tensors_metadata(None) == {}
    with safe_open(GOLD, "pt", "cpu") as f:  # what the reference wrote
        names = set(get_compressed_tensors_metadata(f.metadata()))
    assert names == {"w_bf16", "w_fp16", "w_fp32", "w_fp8", "big_bf16"}


@pytest.mark.gpu
@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_reads_file_written_by_the_reference(device):
    from zipnn_b200 import SafeOpen
    want = make_checkpoint()
    with SafeOpen(GOLD, "pt", device) as f:
        assert set(f.keys()) == set(want)
        for name in f.keys():
            got = f.get_tensor(name)
            assert got.device.type == device
            assert _same(got, want[name]), name
        assert f.get_slice("w_bf16") is NotImplementedError  # as the reference
        assert f.get_slice("ids")[2:4].shape == (2, 100)


@pytest.mark.gpu
def test_writer_output_is_the_reference_file(tmp_path):
    """Our writer on the same checkpoint produces the same compressed entries, byte for byte."""
    from zipnn_b200 import compress_safetensors_file, decompress_safetensors_file
    src = tmp_path / "m.safetensors"
    save_file(make_checkpoint(), str(src), {"format": "pt"})
    path, clen, olen = compress_safetensors_file(str(src))
    assert path.endswith(".znn.safetensors") and clen < olen
    with safe_open(path, "pt", "cpu") as ours, safe_open(GOLD, "pt", "cpu") as ref:
        assert set(ours.keys()) == set(ref.keys())
        assert json.loads(ours.metadata()["znn_compressed_vectors"]) == json.loads(ref.metadata()["znn_compressed_vectors"])
        for name in ref.keys():
            assert torch.equal(ours.get_tensor(name), ref.get_tensor(name)), name
    back = decompress_safetensors_file(path)
    with safe_open(back, "pt", "cpu") as f:
        for name, t in make_checkpoint().items():
            assert _same(f.get_tensor(name), t), name


@pytest.mark.gpu
def test_patch_is_what_a_loader_sees(tmp_path):
    """vLLM's weight iterator does `from safetensors.torch import safe_open` after the patch and
    then `with safe_open(file, framework="pt") as f: for name in f.keys(): f.get_tensor(name)`."""
    import safetensors.torch
    from zipnn_b200 import SafeOpen, compress_safetensors_file, zipnn_safetensors
    src = tmp_path / "m.safetensors"
    want = make_checkpoint()
    save_file(want, str(src))          # no metadata dict in the source file
    path, _, _ = compress_safetensors_file(str(src))
    saved = safetensors.torch.safe_open
    try:
        zipnn_safetensors()
        assert safetensors.torch.safe_open is SafeOpen
        with safetensors.torch.safe_open(path, framework="pt") as f:
            for name in f.keys():
                assert _same(f.get_tensor(name), want[name]), name
    finally:
        safetensors.torch.safe_open = saved


@pytest.mark.gpu
def test_tiny_tensors_survive_the_raw_fallback(tmp_path):
    """The reference stores its in-place-rotated copy when a tensor does not compress and so
    corrupts tiny bf16/fp32 tensors; we store the original bytes."""
    from zipnn_b200 import SafeOpen, compress_safetensors_file
    tensors = {"tiny": torch.tensor([1.5, -2.25, 3.0, 0.1, 7.0, -0.5, 9.0, 1e-3], dtype=torch.bfloat16),
               "big": (torch.randn(4096) * 0.02).to(torch.bfloat16), "one": torch.tensor([3.25], dtype=torch.float32)}
    src = tmp_path / "t.safetensors"
    save_file(tensors, str(src))
    path, _, _ = compress_safetensors_file(str(src))
    with SafeOpen(path, "pt", "cpu") as f:
        assert "tiny" not in f.compressed_tensors_metadata and "big" in f.compressed_tensors_metadata
        for name, t in tensors.items():
            assert _same(f.get_tensor(name), t), name


@pytest.mark.gpu
def test_gpu_load_path_many_tensors_and_deferred_errors(tmp_path):
    """SafeOpen(device="cuda") decodes on side streams without a host sync per tensor: every tensor
    must still be exact when used on the current stream, and a corrupt entry must raise when the
    file is closed."""
    from safetensors.torch import save_file as sf
    from zipnn_b200 import SafeOpen, compress_safetensors_file, load_file
    g = torch.Generator().manual_seed(3)
    want = {}
    for i in range(40):
        dt = (torch.bfloat16, torch.float32, torch.float16)[i % 3]
        n = (1, 17, 4096, 70000, 300001, 1 << 20)[i % 6]
        want[f"t{i}"] = (torch.randn(n, generator=g) * 0.02).to(dt)
    want["ids"] = torch.arange(100, dtype=torch.int64)
    src = tmp_path / "many.safetensors"
    sf(want, str(src))
    path, _, _ = compress_safetensors_file(str(src))
    got = load_file(path, "cuda")
    assert set(got) == set(want)
    for k, t in want.items():
        assert got[k].is_cuda and got[k].dtype == t.dtype and tuple(got[k].shape) == tuple(t.shape)
        assert torch.equal(got[k].cpu().view(torch.uint8), t.view(torch.uint8)), k
    # results are usable immediately on the current stream (no explicit sync by the caller)
    with SafeOpen(path, "pt", "cuda") as f:
        acc = sum(float(f.get_tensor(k).float().sum()) for k in want if k != "ids")
    ref = sum(float(t.float().sum()) for k, t in want.items() if k != "ids")
    assert abs(acc - ref) < 1e-2 * max(1.0, abs(ref))
    # corrupt one compressed entry's type byte inside the file -> error when the file is closed
    import json, struct
    blob = bytearray(open(path, "rb").read())
    hlen = struct.unpack("<Q", blob[:8])[0]
    meta = json.loads(bytes(blob[8: 8 + hlen]))
    a, _ = meta["t5"]["data_offsets"]
    blob[8 + hlen + a + 32 + 1 + 1 + 4 + 0] = 9        # first type byte after the 38-byte header (1-D shape < 2^32)
    bad = tmp_path / "bad.znn.safetensors"
    open(bad, "wb").write(bytes(blob))
    with pytest.raises(RuntimeError, match="corrupt"):
        with SafeOpen(str(bad), "pt", "cuda") as f:
            for k in f.keys():
                f.get_tensor(k)


def test_safetensors_index_matches_the_library(tmp_path):
    """The GPU load path reads compressed entries straight from the file: its own index of
    (offset, length) must address exactly the bytes safetensors hands out."""
    from safetensors.torch import save_file as sf
    from zipnn_b200.safetensors_io import _safetensors_index
    tensors = {"a": torch.arange(1000, dtype=torch.int32), "b": (torch.randn(77, 3) * 0.02).to(torch.bfloat16),
               "c": torch.zeros(0, dtype=torch.float32), "d": torch.randint(0, 255, (4097,), dtype=torch.uint8)}
    path = tmp_path / "idx.safetensors"
    sf(tensors, str(path), {"k": "v"})
    idx = _safetensors_index(str(path))
    blob = open(path, "rb").read()
    assert set(idx) == set(tensors)
    with safe_open(str(path), "pt", "cpu") as f:
        for name in f.keys():
            off, n = idx[name]
            t = f.get_tensor(name)
            assert n == t.numel() * t.element_size()
            assert blob[off: off + n] == t.contiguous().view(torch.uint8).numpy().tobytes(), name


@pytest.mark.gpu
def test_tensor_with_many_general_chunks_is_correct_inside_the_with_block(tmp_path):
    """a tensor with more than 64 chunks that need the general path used to come
    back unwritten until close().  Read it the way vLLM's iterator does -- copy inside the `with` -- and
    compare BEFORE close; both the batched and the per-tensor load path."""
    from zipnn_b200 import SafeOpen, compress_safetensors_file
    g = torch.Generator().manual_seed(5)
    big = (torch.randn(65536 * 150, generator=g) * 0.02).to(torch.bfloat16).to(torch.float32)   # two coded groups per chunk
    small = (torch.randn(4096, generator=g) * 0.02).to(torch.bfloat16)
    src = str(tmp_path / "m.safetensors")
    save_file({"big": big, "small": small}, src)
    path, _, _ = compress_safetensors_file(src)
    for batch in (True, False):
        with SafeOpen(path, "pt", "cuda", batch=batch) as f:
            assert "big" in f.compressed_tensors_metadata
            got_big = f.get_tensor("big").clone()
            got_small = f.get_tensor("small").clone()
            torch.cuda.current_stream().synchronize()
            assert torch.equal(got_big.cpu(), big), f"batch={batch}"
            assert torch.equal(got_small.cpu(), small)


@pytest.mark.gpu
def test_zipnn_hf_loads_znn_checkpoints(tmp_path):
    """zipnn_hf(): transformers' load_state_dict reads `model.safetensors.znn` (the reference's whole-file
    streaming format), and from_pretrained on a directory that only holds the .znn file finds the weights."""
    import transformers
    from transformers import GPT2Config, GPT2LMHeadModel, modeling_utils
    from zipnn_b200 import ZipNN, zipnn_hf
    cfg = GPT2Config(n_layer=2, n_head=2, n_embd=64, vocab_size=257, n_positions=32)
    torch.manual_seed(0)
    model = GPT2LMHeadModel(cfg)
    d = tmp_path / "ckpt"
    model.save_pretrained(str(d), safe_serialization=True)
    plain = d / "model.safetensors"
    raw = plain.read_bytes()
    znn = ZipNN(input_format="byte", bytearray_dtype="float32", is_streaming=True).compress(raw)
    # the whole-file stream is the reference's: frame 0 == the oracle's stream of the first MiB
    from oracle import oracle as O
    import numpy as np
    first = raw[: 1 << 20]
    plan = ZipNN(input_format="byte", bytearray_dtype="float32", is_streaming=True).plan(first)
    want = O.zipnn_compress(plan["header"], np.frombuffer(first, dtype=np.uint8), 4, 1, 220, 262144, 0.95, threads=2)
    assert bytes(znn[: want.size]) == want.tobytes()
    (d / "model.safetensors.znn").write_bytes(bytes(znn))
    plain.unlink()
    zipnn_hf()
    sd = modeling_utils.load_state_dict(str(d / "model.safetensors.znn"))
    ref = model.state_dict()
    assert set(sd) <= set(ref) and all(torch.equal(sd[k].cpu(), ref[k]) for k in sd)
    again = GPT2LMHeadModel.from_pretrained(str(d))
    for k, v in again.state_dict().items():
        assert torch.equal(v, ref[k]), k
    assert (d / "model.safetensors.znn").exists() and not plain.exists()   # replace_local_file defaults to False
    _ = transformers
