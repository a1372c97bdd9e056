"""Probe: does torch symmetric memory give a multicast (NVLS) mapping on this box?  torchrun --nproc-per-node 2"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(1 << 20, dtype=torch.bfloat16, device=dev)
h = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "world", h.world_size, "multicast_ptr", hex(h.multicast_ptr),
      "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "signal_pad_size", h.signal_pad_size, flush=True)
t.fill_(rank + 1)
h.barrier()
torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
torch.cuda.synchronize()
print(rank, "allreduce ->", float(t[0]), float(t[-1]), flush=True)
dist.destroy_process_group()
