#!/usr/bin/env python3
"""NCCL check of zipnn_b200.sharded on real GPUs (run under torchrun, one rank per GPU):
the gathered stream must be byte-identical to the single-GPU stream of the whole tensor, and
scatter + decode must give every rank its slice back.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from zipnn_b200 import ZipNN  # noqa: E402
from zipnn_b200.sharded import ShardedZipNN, byte_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for dt, n_elems in ((torch.bfloat16, 64 * 131072 + 4321), (torch.float32, 40 * 65536 + 12), (torch.float8_e4m3fn, 3_000_001),
                        (torch.bfloat16, 512 * 1024 * 1024)):
        g = torch.Generator(device=dev).manual_seed(99)
        full = (torch.randn(n_elems, generator=g, device=dev) * 0.02).to(dt)
        esz = full.element_size()
        chunk = 131072 if esz == 1 else 262144
        b0, b1 = byte_range(n_elems * esz, chunk, rank, world)
        local_t = full[b0 // esz: b1 // esz].clone()
        z = ShardedZipNN()
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        stream = z.compress(local_t, global_shape=tuple(full.shape), dst=0)
        torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
        if rank == 0:
            want = ZipNN(input_format="torch").compress(full)
            same = stream.numel() == want.numel() and bool(torch.equal(stream, want))
            print(f"[{dt}] n={n_elems}: gathered stream {'==' if same else '!='} single-GPU stream "
                  f"({stream.numel()} bytes), sharded compress+gather {1e3 * (t1 - t0):.1f} ms", flush=True)
            ok &= same
        back = z.decompress(stream if rank == 0 else None, src=0, device=dev)
        good = back.dtype == dt and bool(torch.equal(back.view(torch.uint8), local_t.view(torch.uint8)))
        flag = torch.tensor([1 if good else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"[{dt}] scatter + decode: {'ok' if flag.item() else 'MISMATCH'}", flush=True)
            ok &= bool(flag.item())
        del full, local_t, stream, back
        torch.cuda.empty_cache()
    dist.barrier()
    if rank == 0:
        print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
