"""Probe: does torch symmetric memory give peer-mapped pointers on this box? (run under torchrun)"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm_mem.empty(64, dtype=torch.float64, device=torch.device("cuda", local))
t.zero_()
h = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in h.signal_pad_ptrs][:2], flush=True)
# write my rank into slot[rank] of every peer through the mapped pointers
for p in range(world):
    peer = h.get_buffer(p, (64,), torch.float64)
    peer[rank] = float(rank + 1)
torch.cuda.synchronize()
dist.barrier()
print(rank, "local view", t[:world].tolist(), flush=True)
dist.destroy_process_group()
