import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); lr=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = sm.empty(1<<20, dtype=torch.float64, device=torch.device("cuda", lr))
h = sm.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in h.buffer_ptrs], "mc", hex(h.multicast_ptr) if h.multicast_ptr else None, "world", h.world_size, flush=True)
t.fill_(float(rank+1))
h.barrier()
peer = h.get_buffer((rank+1)%world, (1<<20,), torch.float64)
print(rank, "peer value", float(peer[5]), flush=True)
h.barrier()
dist.destroy_process_group()
