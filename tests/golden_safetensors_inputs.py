import torch


def make_checkpoint():
    """Small synthetic checkpoint: the reference's test pattern per dtype, large enough that
    every float tensor compresses (the reference corrupts tensors it leaves raw)."""
    g = torch.Generator().manual_seed(4321)
    out = {}
    for name, dt in (("w_bf16", torch.bfloat16), ("w_fp16", torch.float16), ("w_fp32", torch.float32),
                     ("w_fp8", torch.float8_e4m3fn)):
        t = torch.randn(100, 100, generator=g)
        t[:50] = 42.0
        out[name] = t.to(dt)
    out["big_bf16"] = (torch.randn(300, 257, generator=g) * 0.02).to(torch.bfloat16)
    out["ids"] = torch.arange(1000, dtype=torch.int64).reshape(10, 100)
    return out
