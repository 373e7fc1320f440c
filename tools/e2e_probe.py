"""Breakdown of the host-buffer round trip: raw PCIe copy rates, compress_host, decompress_host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_tensor
from zipnn_b200 import ZipNN

if os.environ.get("ZIPNN_SLAB_MIB"):      # experiment knob: slab size of the host pipelines (read by the library)
    os.environ["ZIPNN_B200_HOST_SLAB_BYTES"] = str(int(os.environ["ZIPNN_SLAB_MIB"]) << 20)

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
n = int(gib * (1 << 30))
dev = torch.device("cuda", 0)
t = make_tensor(n, torch.bfloat16, dev, 1234)
ht = torch.empty(t.numel(), dtype=torch.bfloat16, pin_memory=True)
ht.copy_(t)
hs = torch.empty(n + (n >> 6) + 4096, dtype=torch.uint8, pin_memory=True)
hd = torch.empty(t.numel(), dtype=torch.bfloat16, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device=dev)

def tm(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

h2d = tm(lambda: d.copy_(ht.view(torch.uint8), non_blocking=True))
d2h = tm(lambda: hd.view(torch.uint8).copy_(d, non_blocking=True))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty(n, dtype=torch.uint8, device=dev)
def both():
    with torch.cuda.stream(s1):
        d.copy_(ht.view(torch.uint8), non_blocking=True)
    with torch.cuda.stream(s2):
        hd.view(torch.uint8).copy_(d2, non_blocking=True)
dup = tm(both)
print(f"H2D {n/h2d/1e9:.1f} GB/s  D2H {n/d2h/1e9:.1f} GB/s  duplex {2*n/dup/1e9:.1f} GB/s (sum)", flush=True)
zs = ZipNN(input_format="torch").compress(ht, out=hs)
c = len(zs)
tc = tm(lambda: ZipNN(input_format="torch").compress(ht, out=hs))
td = tm(lambda: ZipNN(input_format="torch").decompress(zs, out=hd))
print(f"compress_host {tc*1e3:.0f} ms ({n/tc/1e9:.1f} GB/s; copies alone {((n/ (n/h2d)) and (h2d + c/n*d2h))*1e3:.0f} ms)  "
      f"decompress_host {td*1e3:.0f} ms ({n/td/1e9:.1f} GB/s; copies alone {(c/n*h2d + d2h)*1e3:.0f} ms, overlapped {max(c/n*h2d, d2h)*1e3:.0f} ms)", flush=True)
