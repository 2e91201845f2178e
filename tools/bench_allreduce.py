#!/usr/bin/env python
"""Standalone timing of the gradient-bucket all-reduce: libsgan's copy-engine path vs NCCL, per bucket size.
    python -m torch.distributed.run --nproc-per-node N tools/bench_allreduce.py"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


class _Store:
    pass


def main():
    runtime = importlib.import_module("scrabble-gan_b200.runtime")
    dp = importlib.import_module("scrabble-gan_b200.dp")
    rt = runtime.Runtime(device=int(os.environ.get("LOCAL_RANK", "0")), mode="bf16")
    runtime.set_runtime(rt)
    dp.init_data_parallel(rt)
    for mb in (22, 65, 150):
        n = mb * (1 << 20) // 4
        st = _Store()
        st.g = torch.randn(n, device=rt.device)
        ref = st.g.clone()
        res = {}
        for name in ("ce", "sm", "nccl"):
            def one():
                if name in ("ce", "sm"):
                    rt.allreduce_async_(st.g, store=st, exposed=(name == "sm")).wait()
                else:
                    dist.all_reduce(ref)
            for _ in range(3):
                one()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                one()
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 10
        if rt.rank == 0:
            print("bucket {:4d} MB, world {}: copy engines {:.3f} ms   SM pull kernels {:.3f} ms   NCCL {:.3f} ms".format(
                mb, rt.world_size, res["ce"], res["sm"], res["nccl"]), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
