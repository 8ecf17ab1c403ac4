"""Two or more ranks (torchrun, one per GPU): sgb_gae_allgather == sgb_gae + NCCL all_gather_into_tensor, bit for bit, in
every rank's gather buffers.   torchrun --nproc-per-node 2 tests/tools/fused_gather_check.py [0|1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

from sigmarl_b200.rollout import RolloutBuffer, all_gather_advantages, compute_gae, gae_allgather

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
mc = {"auto": None, "1": True, "0": False}[sys.argv[1] if len(sys.argv) > 1 else "auto"]
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for T, B, N in ((16, 1031, 6), (24, 515, 8)):           # one column per thread / four columns per thread (16-byte stores)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    buf = RolloutBuffer(T, B, N, 4, dev, world=world, rank=rank, symmetric=True)
    for name in ("reward", "value", "next_value"):
        getattr(buf, name).copy_(torch.randn(T, B, N, device=dev, generator=g))
    buf.done.copy_((torch.rand(T, B, device=dev, generator=g) < 0.1).to(torch.uint8))
    for rep in range(3):                                   # repeated: the barriers must order successive rounds
        buf.reward.add_(1.0)
        compute_gae(buf, 0.99, 0.9)
        all_gather_advantages(buf)
        torch.cuda.synchronize()
        want_a, want_t = buf.adv_all.clone(), buf.target_all.clone()
        assert len({float(want_a[w].sum()) for w in range(world)}) == world      # the ranks hold different data
        buf.adv_all.zero_(); buf.target_all.zero_()
        gae_allgather(buf, 0.99, 0.9, multicast=mc)
        torch.cuda.synchronize()
        assert torch.equal(buf.adv_all, want_a) and torch.equal(buf.target_all, want_t), f"rank {rank} round {rep} N {N}"
print(f"FUSED-GATHER-OK rank {rank} world {world} multicast {'yes' if mc else 'no'}", flush=True)
dist.barrier()
dist.destroy_process_group()
