"""BASELINE.json configs[4]: throughput sweep over tokens / GPU / step at N GPUs, uniform and Zipf token ids.
    python tools/scale_sweep.py                                   (1 GPU)
    torchrun --nproc-per-node N tools/scale_sweep.py              (N GPUs, weak scaling: tables replicated, batch sharded)
MoT-sum (runs/71) at 768 = 16 x 48, bf16: one step = fused forward + backward (+ the exchange of the gradient bucket for
N > 1), device-resident inputs, CUDA-event timed, max over ranks.  Prints one markdown row per size."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "mixture-of-tokenizers_b200")):
    sys.path.insert(0, p_)
import torch
import torch.distributed as dist
import mot_b200
from mot_b200 import ops, dp
import bench

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
Dt, bd, bpt, V = 768, 48, 16, 50257
sizes = [int(x) for x in os.environ.get("SWEEP_SIZES", "1024,4096,16384,65536,262144,1048576").split(",")]
g = torch.Generator(device=dev).manual_seed(12345)
E_tok = torch.randn(V, Dt, generator=g, device=dev).bfloat16()
E_byte = torch.randn(458, bd, generator=g, device=dev).bfloat16()
bucket = dp.GradBucket([torch.nn.Parameter(E_tok, requires_grad=False), torch.nn.Parameter(E_byte, requires_grad=False)],
                       symmetric="auto" if world > 1 else False)
gt, gb = bucket.views()
spec = mot_b200.MixSpec(combine="add", slot_major=True)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


if rank == 0:
    print(f"| tokens/GPU | dist | GPUs | exchange | step us | compute-only us | tokens/s (all GPUs) | per GPU vs the 1-GPU compute rate | touched rows/rank |")
    print("|---|---|---|---|---|---|---|---|---|")
for N in sizes:
    for dname in ("uniform", "zipf"):
        tok = bench.make_tokens(N, dname, 12345 + rank).to(dev)
        gd = torch.Generator(device=dev).manual_seed(7 + rank)
        ids = torch.randint(0, 458, (bpt, N), generator=gd, device=dev, dtype=torch.int32)
        gout = torch.randn(N, Dt, generator=gd, device=dev).bfloat16()
        out = torch.empty(N, Dt, dtype=torch.bfloat16, device=dev)
        desc = ops.make_desc(spec, N, E_tok, E_byte, bpt, ids=ids, ttb=None, has_lam=False, seq_len=N)
        saved = ops.embed_bwd_uses_saved(desc)
        rstd = torch.empty(N, dtype=torch.float32, device=dev) if saved else None
        ws = ops.acquire_workspace(desc, dev)

        def step(exchange=True):
            ops.embed_plan_async(desc, tok, ws, dev)
            ops.embed_forward_out(desc, tok, ids, None, E_tok, E_byte, None, out, rstd=rstd)
            ops.embed_plan_join(ws, dev)
            ops.embed_backward_out(desc, tok, ids, None, E_tok, E_byte, None, gout, gt, gb, None, ws.buf, plan_ready=True,
                                   ws_clean=True, out_saved=out if saved else None, rstd=rstd, plan_joined=True)
            ws.clean = True
            if world > 1 and exchange:
                bucket.mark_rows(desc, ws.buf)
                bucket.all_reduce_avg()

        def timed(fn, reps):
            for _ in range(5):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b) / reps * 1e3], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        reps = 100 if N <= 65536 else (30 if N <= 262144 else 10)
        us = timed(step, reps)
        us_c = timed(lambda: step(False), reps) if world > 1 else us
        dens = float(torch.unique(tok).numel()) / V
        if rank == 0:
            print(f"| {N} | {dname} | {world} | {bucket.algo if world > 1 else '-'} | {us:.1f} | {us_c:.1f} | {world * N / us:.1f} M | "
                  f"{us_c / us:.2f} | {dens:.2f} |", flush=True)
        ops.release_workspace(ws)
        del ws
if world > 1:
    dist.destroy_process_group()
