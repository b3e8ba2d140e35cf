"""Time of the gradient exchange alone: dist.all_reduce of the 94 MB flat fp32 gradient (and of its four buckets) on the
GPUs of one box, CUDA events, max over ranks.  torchrun --nproc-per-node N tools/nccl_allreduce_bench.py"""
import os
import sys
import torch
import torch.distributed as dist


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sys.stdout.flush()
    fd = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    n = 23510081
    flat = torch.randn(n, device=dev)
    res = {}
    for name, sizes in (("one_94MB", [n]), ("buckets_60_28_5_1", [14964736, 7098368, 1219584, 227393])):
        chunks, off = [], 0
        for sz in sizes:
            chunks.append(flat[off:off + sz])
            off += sz
        for _ in range(5):
            for c in chunks:
                dist.all_reduce(c)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        a.record()
        for _ in range(reps):
            for c in chunks:
                dist.all_reduce(c)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = round(float(t), 4)
    if rank == 0:
        tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("NCCL_") and k not in ("NCCL_DEBUG",))
        os.write(fd, (f"world {dist.get_world_size()} [{tag}] ms per exchange: {res}\n").encode())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
