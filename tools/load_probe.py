"""cProfile of the GPU load path on a synthetic gpt2-shaped .znn.safetensors in /dev/shm."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safetensors.torch import save_file
from tools.model_bench import MODELS
from zipnn_b200 import SafeOpen, compress_safetensors_file

model = sys.argv[1] if len(sys.argv) > 1 else "gpt2"
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 0
shapes, dtype = MODELS[model](layers)
g = torch.Generator(device="cuda").manual_seed(1)
tensors = {}
for name, shp in shapes.items():
    n = 1
    for d in shp:
        n *= d
    tensors[name] = (torch.randn(n, generator=g, device="cuda") * 0.02).to(dtype).reshape(shp).cpu()
src = f"/dev/shm/{model}_probe.safetensors"
save_file(tensors, src)
path, _, _ = compress_safetensors_file(src)

def load():
    with SafeOpen(path, "pt", "cuda") as f:
        out = {k: f.get_tensor(k) for k in f.keys()}
    torch.cuda.synchronize()
    return out

for i in range(3):
    t0 = time.perf_counter(); load(); print("load", i, round(time.perf_counter() - t0, 3), flush=True)
pr = cProfile.Profile(); pr.enable(); load(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
os.remove(src); os.remove(path)
