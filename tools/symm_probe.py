"""Development probe: does torch symmetric memory (peer-mapped buffers + signal pads) work on this box?"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", lr))
t.fill_(float(rank + 1))
h = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in h.buffer_ptrs], "pads", [hex(p) for p in h.signal_pad_ptrs], "pad size", h.signal_pad_size,
      "multicast", h.has_multicast_support, hex(h.multicast_ptr) if h.has_multicast_support else None, flush=True)
h.barrier()
peer = (rank + 1) % world
pb = h.get_buffer(peer, (1 << 20,), torch.float32)
x = pb[:8].clone()
torch.cuda.synchronize()
print(rank, "read from peer", peer, x.tolist(), flush=True)
h.barrier()
pb[8:16] = 100.0 + rank
h.barrier()
torch.cuda.synchronize()
print(rank, "my buffer after peer write", t[6:18].tolist(), flush=True)
# bandwidth of a peer read
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
big = symm_mem.empty(64 << 20, dtype=torch.float32, device=torch.device("cuda", lr))
hb = symm_mem.rendezvous(big, dist.group.WORLD)
src = hb.get_buffer(peer, (64 << 20,), torch.float32)
dst = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
hb.barrier()
for _ in range(3):
    dst.copy_(src)
e0.record()
for _ in range(10):
    dst.copy_(src)
e1.record()
torch.cuda.synchronize()
print(rank, "peer read GB/s", 10 * 256e6 / (e0.elapsed_time(e1) * 1e-3) / 1e9, flush=True)
hb.barrier()
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
