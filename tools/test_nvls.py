"""N-rank check + timing of the library's NVLS all-reduce against NCCL.  torchrun --nproc-per-node N tools/test_nvls.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200"))
import torch, torch.distributed as dist
from mot_b200 import dp
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
for dtype, shapes in [(torch.bfloat16, [(50257, 768), (458, 48)]), (torch.float32, [(1000, 40), (14, 8)])]:
    params = [torch.nn.Parameter(torch.empty(s, dtype=dtype, device=dev), requires_grad=False) for s in shapes]
    b = dp.GradBucket(params, symmetric=True)
    assert b._symm is not None, "no multicast mapping"
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    src = torch.randn(b.flat.numel(), generator=g, device=dev).to(dtype)
    for trial in range(3):
        b.flat.copy_(src)
        ref = src.float()                      # fp32 reference (NCCL's bf16 all-reduce rounds at every hop)
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        ref /= world
        b.all_reduce_avg()
        torch.cuda.synchronize()
        err = float((b.flat.double() - ref.double()).abs().max() / ref.double().abs().max())
        assert err <= (2.0 ** -7 if dtype == torch.bfloat16 else 1e-6), (rank, trial, err)   # bf16: switch sum rounded, then the scaled value rounded
    if rank == 0:
        print(f"{dtype} n={b.flat.numel()} ok, err {err:.2e}", flush=True)
    def timeit(fn, reps=20):
        for _ in range(3): fn()
        dist.barrier(); torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        e.record(); torch.cuda.synchronize()
        return a.elapsed_time(e) / reps * 1e3
    t_own = timeit(b.all_reduce_avg)
    if dtype == torch.bfloat16 and os.environ.get("MOT_AR_SWEEP"):
        for blocks, threads, unroll in [(8, 1024, 8), (16, 1024, 8), (24, 1024, 8), (36, 1024, 8), (16, 1024, 4), (24, 512, 8), (36, 512, 8)]:
            os.environ.update(MOT_AR_BLOCKS=str(blocks), MOT_AR_THREADS=str(threads), MOT_AR_UNROLL=str(unroll))
            t = timeit(b.all_reduce_avg)
            if rank == 0:
                print(f"  sweep world {world}: blocks {blocks} threads {threads} unroll {unroll}: {t:.1f} us", flush=True)
        for k in ("MOT_AR_BLOCKS", "MOT_AR_THREADS", "MOT_AR_UNROLL"):
            os.environ.pop(k, None)
    ref = src.clone()
    t_nccl = timeit(lambda: dist.all_reduce(ref, op=dist.ReduceOp.AVG))
    if rank == 0:
        mb = b.flat.numel() * b.flat.element_size() / 1e6
        print(f"world {world} {mb:.1f} MB: own NVLS kernel {t_own:.1f} us ({mb / t_own * 1e3:.0f} GB/s algbw) | NCCL {t_nccl:.1f} us", flush=True)
dist.destroy_process_group()
