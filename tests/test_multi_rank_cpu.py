"""CPU, world_size 2 over gloo: the N>1 host logic — env sharding by index and the single collective
(in-place all-gather of the advantage / value-target buffers)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gae_numpy(reward, value, next_value, done, gamma, lmbda):
    """Plain restatement of the GAE recurrence for checking."""
    T = reward.shape[0]
    adv = np.zeros_like(reward)
    a = np.zeros_like(reward[0])
    for t in range(T - 1, -1, -1):
        nd = (1.0 - done[t].astype(np.float32))[:, None]
        delta = reward[t] + np.float32(gamma) * next_value[t] * nd - value[t]
        a = delta + np.float32(gamma) * np.float32(lmbda) * nd * a
        adv[t] = a
    return adv, adv + value


def _worker(rank, world, port, out):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sigmarl_b200.rollout import RolloutBuffer, all_gather_advantages, shard_range
    T, B_total, N, D = 5, 8, 3, 4
    off, B = shard_range(B_total, rank, world)
    rng = np.random.default_rng(0)
    full = {k: rng.standard_normal((T, B_total, N)).astype(np.float32) for k in ("reward", "value", "next_value")}
    done = rng.random((T, B_total)) < 0.2
    adv, tgt = gae_numpy(full["reward"], full["value"], full["next_value"], done, 0.99, 0.9)
    buf = RolloutBuffer(T, B, N, D, "cpu", world=world, rank=rank)
    # each rank fills its own slot from ITS env shard (on a GPU this is what sgb_gae writes)
    buf.advantage.copy_(torch.from_numpy(adv[:, off:off + B]))
    buf.value_target.copy_(torch.from_numpy(tgt[:, off:off + B]))
    a_all, t_all = all_gather_advantages(buf)
    got_a = torch.cat([a_all[r] for r in range(world)], dim=1).numpy()
    got_t = torch.cat([t_all[r] for r in range(world)], dim=1).numpy()
    ok = np.array_equal(got_a, adv) and np.array_equal(got_t, tgt) and (off, B) == (rank * 4, 4)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_env_sharding_and_all_gather_world2():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        port = 29500 + (os.getpid() % 2000)
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_shard_range_rejects_uneven_split():
    sys.path.insert(0, REPO)
    from sigmarl_b200.rollout import shard_range
    import pytest
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
    assert shard_range(262144, 3, 8) == (3 * 32768, 32768)
