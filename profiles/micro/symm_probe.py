"""Probe: torch symmetric memory (peer pointers / NVLS multicast) on this box.
   torchrun --nproc-per-node N profiles/micro/symm_probe.py"""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 64 << 20
t = symm_mem.empty(world * n, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast ptr", hex(hdl.multicast_ptr or 0), "signal pad", hdl.signal_pad_size, flush=True)
# push own slot to every peer with plain copies through the peer views, then barrier
src = torch.full((n,), float(rank + 1), device=dev)
hdl.barrier(channel=0)
torch.cuda.synchronize()
for it in range(3):
    t0 = time.perf_counter()
    for w in range(world):
        peer = hdl.get_buffer((rank + w) % world, (world * n,), torch.float32)
        peer[rank * n:(rank + 1) * n].copy_(src)
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(rank, f"push all-gather {world * n * 4 / 1e6:.0f} MB total per rank in {dt * 1e3:.3f} ms -> {(world - 1) * n * 4 / dt / 1e9:.1f} GB/s received", flush=True)
ok = all(bool((t[w * n:(w + 1) * n] == w + 1).all()) for w in range(world))
# NCCL for comparison
full = torch.empty(world * n, device=dev); 
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dist.all_gather_into_tensor(full, src)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(rank, f"nccl all-gather {dt * 1e3:.3f} ms -> {(world - 1) * n * 4 / dt / 1e9:.1f} GB/s received", flush=True)
print(rank, "content ok", ok, flush=True)
dist.destroy_process_group()
