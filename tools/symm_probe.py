"""Probe (2+ GPUs, torchrun): does torch symmetric memory give usable NVLink peer pointers here?  Each rank writes a pattern into
every peer's buffer with a plain CUDA copy through the peer pointer, barriers, and checks what the peers wrote into its own."""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 64 << 20  # 64 MiB of bf16 per slot
buf = symm.empty((world, n), dtype=torch.bfloat16, device=dev)
hdl = symm.rendezvous(buf, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hdl.has_multicast_support, flush=True)
src = torch.full((n,), float(rank + 1), dtype=torch.bfloat16, device=dev)
hdl.barrier(channel=0)
for peer in range(world):
    remote = hdl.get_buffer(peer, (world, n), torch.bfloat16)
    remote[rank].copy_(src)   # store into the peer's slot [rank]
hdl.barrier(channel=0)
torch.cuda.synchronize()
ok = all(float(buf[r, 0]) == r + 1 and float(buf[r, -1]) == r + 1 for r in range(world))
# bandwidth of a peer store
peer = (rank + 1) % world
remote = hdl.get_buffer(peer, (world, n), torch.bfloat16)
for _ in range(3): remote[rank].copy_(src)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): remote[rank].copy_(src)
e1.record(); torch.cuda.synchronize()
gbs = 10 * n * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9
# barrier latency
hdl.barrier(channel=0); torch.cuda.synchronize()
e0.record()
for _ in range(50): hdl.barrier(channel=0)
e1.record(); torch.cuda.synchronize()
print(f"rank {rank}: ok={ok} peer-store {gbs:.0f} GB/s, barrier {e0.elapsed_time(e1) / 50 * 1e3:.1f} us", flush=True)
dist.destroy_process_group()
