"""Timing of the exchange kernels against NCCL at N ranks (torchrun --nproc-per-node N tools/dp_bench.py):
whole-bucket exchange per algorithm and launch shape, the bucket as 4 ranges, and NCCL's all-reduce."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200"))
import torch
import torch.distributed as dist
from mot_b200 import dp

rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
Dt = int(os.environ.get("DP_BENCH_DT", "768"))
shapes = [(50257, Dt), (458, Dt // 16)]


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(e) / reps * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def say(*a):
    if rank == 0:
        print(*a, flush=True)


for algo in ("p2p", "nvls"):
    os.environ["MOT_DP_ALGO"] = algo
    params = [torch.nn.Parameter(torch.empty(s, dtype=torch.bfloat16, device=dev), requires_grad=False) for s in shapes]
    b = dp.GradBucket(params, symmetric=True, n_slabs=4)
    if b._symm is None or b.algo != algo:
        say(f"{algo}: unavailable (algo {b.algo})")
        continue
    mb = b.flat.numel() * 2 / 1e6
    cfgs = [(None, None, None)]
    if algo == "p2p":
        cfgs += [(bl, 512, u) for bl in (8, 16, 32, 64, 96) for u in (2, 4)]
    else:
        cfgs += [(bl, th, u) for bl, th, u in ((4, 1024, 8), (8, 1024, 8), (16, 1024, 8), (8, 1024, 4), (16, 512, 16), (32, 512, 16), (8, 512, 16))]
    for bl, th, u in cfgs:
        for k, v in (("MOT_AR_BLOCKS", bl), ("MOT_AR_THREADS", th), ("MOT_AR_UNROLL", u)):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        t = timeit(b.all_reduce_avg)
        say(f"world {world} {algo} {mb:.1f} MB blocks={bl} threads={th} unroll={u}: {t:.1f} us ({mb / t * 1e3:.0f} GB/s algbw)")
    for k in ("MOT_AR_BLOCKS", "MOT_AR_THREADS", "MOT_AR_UNROLL"):
        os.environ.pop(k, None)
    n = b.flat.numel()
    cuts = [n * i // 4 // 8 * 8 for i in range(4)] + [n]

    def ranges():
        for i in range(4):
            b.exchange_async(cuts[i], cuts[i + 1], last=(i == 3))
        b.wait()
    say(f"world {world} {algo} as 4 ranges on the exchange stream: {timeit(ranges):.1f} us")
    small = dp.GradBucket([torch.nn.Parameter(torch.empty(4096, 8, dtype=torch.bfloat16, device=dev), requires_grad=False)], symmetric=True)
    say(f"world {world} {algo} 64 KB (latency): {timeit(small.all_reduce_avg):.1f} us")
buf = torch.empty(sum(a * c for a, c in shapes), dtype=torch.bfloat16, device=dev)
say(f"world {world} NCCL all_reduce(AVG) {buf.numel() * 2 / 1e6:.1f} MB: {timeit(lambda: dist.all_reduce(buf, op=dist.ReduceOp.AVG)):.1f} us")
dist.destroy_process_group()
