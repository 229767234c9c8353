"""Host -> device copy ceiling of the end-to-end path: every rank copies its own pinned image batch to its GPU, nothing else.

    python scripts/h2d_bw.py                     # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 scripts/h2d_bw.py

Prints one JSON line: per-rank and aggregate GB/s for the bf16 batch (308 MB) and the uint8 batch (154 MB) of bench.py's e2e leg,
and the images/s ceiling each implies (copy time only)."""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = 1024
    out = {"n_gpus": world, "batch": B}
    for name, dt in (("bf16", torch.bfloat16), ("uint8", torch.uint8)):
        host = [torch.zeros(B, 3, 224, 224, dtype=dt).pin_memory() for _ in range(2)]
        dst = [torch.empty(B, 3, 224, 224, dtype=dt, device=dev) for _ in range(2)]
        for i in range(3):
            dst[i & 1].copy_(host[i & 1], non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for i in range(n):
            dst[i & 1].copy_(host[i & 1], non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        nbytes = host[0].numel() * host[0].element_size()
        out[name] = {"bytes": nbytes, "ms_max_over_ranks": float(ms), "gbs_per_gpu": nbytes / float(ms) / 1e6,
                     "gbs_total": world * nbytes / float(ms) / 1e6, "img_s_ceiling": world * B / (float(ms) / 1e3)}
        del host, dst
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
