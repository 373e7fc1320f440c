#!/usr/bin/env python3
"""Model-shaped load-path benchmark (BASELINE.json configs 1-3).

No network: the checkpoints are synthesised with the real tensor shapes and randn*0.02 values
(SURVEY.md section 8d), written as .safetensors to --dir (default /dev/shm), compressed to
.znn.safetensors with zipnn_b200.compress_safetensors_file, and then loaded the way vLLM's weight
iterator does it (`with safe_open(f, framework="pt") as f: for name in f.keys(): f.get_tensor(name)`)
through zipnn_b200.SafeOpen(device="cuda") -- compressed bytes H2D, decode on the GPU.

usage: python tools/model_bench.py [gpt2|llama3-8b|granite-8b] [--layers N] [--dir /dev/shm]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from safetensors import safe_open  # noqa: E402
from safetensors.torch import save_file  # noqa: E402

from zipnn_b200 import SafeOpen, compress_safetensors_file  # noqa: E402


def gpt2_shapes(layers=12):
    h = 768
    s = {"wte.weight": (50257, h), "wpe.weight": (1024, h), "ln_f.weight": (h,), "ln_f.bias": (h,)}
    for i in range(layers):
        p = f"h.{i}."
        s.update({p + "ln_1.weight": (h,), p + "ln_1.bias": (h,), p + "attn.c_attn.weight": (h, 3 * h),
                  p + "attn.c_attn.bias": (3 * h,), p + "attn.c_proj.weight": (h, h), p + "attn.c_proj.bias": (h,),
                  p + "ln_2.weight": (h,), p + "ln_2.bias": (h,), p + "mlp.c_fc.weight": (h, 4 * h),
                  p + "mlp.c_fc.bias": (4 * h,), p + "mlp.c_proj.weight": (4 * h, h), p + "mlp.c_proj.bias": (h,)})
    return s, torch.float32


def llama_like(layers, h, ffn, vocab, kv):
    s = {"model.embed_tokens.weight": (vocab, h), "model.norm.weight": (h,), "lm_head.weight": (vocab, h)}
    for i in range(layers):
        p = f"model.layers.{i}."
        s.update({p + "self_attn.q_proj.weight": (h, h), p + "self_attn.k_proj.weight": (kv, h),
                  p + "self_attn.v_proj.weight": (kv, h), p + "self_attn.o_proj.weight": (h, h),
                  p + "mlp.gate_proj.weight": (ffn, h), p + "mlp.up_proj.weight": (ffn, h),
                  p + "mlp.down_proj.weight": (h, ffn), p + "input_layernorm.weight": (h,),
                  p + "post_attention_layernorm.weight": (h,)})
    return s


MODELS = {
    "gpt2": lambda L: gpt2_shapes(L or 12),
    "llama3-8b": lambda L: (llama_like(L or 32, 4096, 14336, 128256, 1024), torch.bfloat16),
    "granite-8b": lambda L: (llama_like(L or 40, 4096, 12800, 49155, 1024), torch.float16),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model", nargs="?", default="gpt2", choices=sorted(MODELS))
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--keep", action="store_true")
    args = ap.parse_args()
    shapes, dtype = MODELS[args.model](args.layers)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1234)
    tensors = {}
    for name, shp in shapes.items():
        n = 1
        for d in shp:
            n *= d
        tensors[name] = (torch.randn(n, generator=g, device=dev) * 0.02).to(dtype).reshape(shp).cpu()
    nbytes = sum(t.numel() * t.element_size() for t in tensors.values())
    src = os.path.join(args.dir, f"{args.model}.safetensors")
    save_file(tensors, src, {"format": "pt"})
    t0 = time.perf_counter()
    path, clen, olen = compress_safetensors_file(src)
    torch.cuda.synchronize()
    t_comp = time.perf_counter() - t0

    def load(opener, p):
        t0 = time.perf_counter()
        got = {}
        with opener(p, "pt", "cuda") as f:
            for name in f.keys():
                got[name] = f.get_tensor(name)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, got

    load(SafeOpen, path)                       # warm-up (page cache, allocator, kernels)
    t_znn, got = load(SafeOpen, path)
    ok = all(torch.equal(got[k].view(torch.uint8).cpu(), tensors[k].view(torch.uint8)) for k in tensors)
    del got
    load(safe_open, src)
    t_raw, _ = load(safe_open, src)
    row = dict(model=args.model, dtype=str(dtype).replace("torch.", ""), tensors=len(tensors), bytes=nbytes,
               file_ratio=round(os.path.getsize(path) / os.path.getsize(src), 4), exact=ok,
               compress_file_s=round(t_comp, 3), compress_file_gbs=round(nbytes / t_comp / 1e9, 2),
               load_znn_to_cuda_s=round(t_znn, 3), load_znn_gbs=round(nbytes / t_znn / 1e9, 2),
               load_plain_safetensors_to_cuda_s=round(t_raw, 3), load_plain_gbs=round(nbytes / t_raw / 1e9, 2))
    print(json.dumps(row), flush=True)
    if not args.keep:
        os.remove(src)
        os.remove(path)


if __name__ == "__main__":
    main()
