"""Rollout collection + GAE + the one collective of the design ("next" row).

Replaces, for this path, ``SyncDataCollectorCustom.rollout`` (``helper_training.py:686-788``) and the TorchRL
GAE configured in ``optimization_module.py:62-67`` / ``mappo_cavs.py:342-378``:

* ``RolloutBuffer`` keeps ``[T, B, N, ...]`` tensors on the device, written in place each step (no per-step
  TensorDict stacking).
* ``collect`` drives ``RoadTrafficEnv`` for T steps with a policy callable, with masked device resets.
* ``compute_gae`` runs the reverse-scan CUDA kernel (``sgb_gae``) and writes advantage / value target
  DIRECTLY into this rank's slot of the all-gather buffers, so the collective that follows needs no copy.
* ``all_gather_advantages`` is the single data exchange of the multi-GPU design: envs are sharded by index
  over ranks with no communication during the rollout; at PPO-update time every rank gathers everyone's
  ``[T, B_local, N]`` advantage and value-target buffers (NCCL ``all_gather_into_tensor``, in place).
* ``gae_allgather`` is the two of them as ONE kernel (``sgb_gae_allgather``): with the gather buffers in symmetric
  (peer-mapped) memory — ``RolloutBuffer(symmetric=True)`` — every advantage / value target is stored straight into all
  ranks' buffers over NVLink (or once to the NVSwitch multicast address) while the scan runs.
"""
import ctypes as C
from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import lib as _lib


def shard_range(num_envs_total: int, rank: int, world: int):
    """Contiguous env-index range owned by `rank` (envs are independent: SURVEY.md A.7)."""
    if num_envs_total % world:
        raise ValueError(f"num_envs_total={num_envs_total} must be divisible by world size {world}")
    per = num_envs_total // world
    return rank * per, per


class RolloutBuffer:
    def __init__(self, T: int, B: int, N: int, D: int, device, world: int = 1, rank: int = 0, symmetric: bool = False,
                 group=None):
        """symmetric: allocate the gather buffers in torch symmetric memory and map them into every rank of `group`
        (one process per GPU of one node, NCCL process group initialised): needed by ``gae_allgather``."""
        z = lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype, device=device)  # noqa: E731
        self.T, self.B, self.N, self.D, self.world, self.rank = T, B, N, D, world, rank
        self.symm = None
        self.obs = z(T, B, N, D)
        self.action = z(T, B, N, 2)
        self.reward = z(T, B, N)
        self.done = z(T, B, dtype=torch.uint8)
        self.value = z(T, B, N)
        self.next_value = z(T, B, N)
        # all-gather buffers [world, T, B, N]; this rank's GAE output is written into slot `rank`
        if symmetric and world > 1:
            import torch.distributed._symmetric_memory as symm_mem
            both = symm_mem.empty((2, world, T, B, N), dtype=torch.float32, device=device)
            both.zero_()
            self.symm = symm_mem.rendezvous(both, group if group is not None else dist.group.WORLD)
            self._both = both
            self.adv_all, self.target_all = both[0], both[1]
        else:
            self.adv_all = z(world, T, B, N)
            self.target_all = z(world, T, B, N)

    @property
    def advantage(self):
        return self.adv_all[self.rank]

    @property
    def value_target(self):
        return self.target_all[self.rank]


def collect(env, policy: Callable[[torch.Tensor], torch.Tensor], buf: RolloutBuffer,
            value_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, in_place: bool = True):
    """T environment steps.  `policy(obs[B,N,D]) -> action[B,N,2]`; `value_fn(obs) -> [B,N]` (optional).

    TorchRL semantics kept (SURVEY.md assumption A5): the stored next-state value of a done env is the value of
    its step-time observation (before the reset); the observation the policy sees next is the post-reset one.

    in_place: the env's outputs are bound to the rollout buffer (``RoadTrafficEnv.bind``): step t writes
    ``buf.reward[t]``, ``buf.done[t]`` and — as the next step's input — ``buf.obs[t + 1]`` directly; no per-step
    stacking or copies (the reference stacks TensorDicts, ``helper_training.py:745-770``).  in_place=False keeps the
    env's own buffers and copies (same results).  In both modes ``buf.action[t]`` holds the policy's RAW output: the
    step kernel clamps its own copy in place (``helper_training.py:807-818`` clamps ``agent.action.u``, not the
    collector's tensordict), so PPO evaluates log-probabilities on the action that was sampled."""
    T = buf.T
    if not in_place:
        obs = env.obs
        for t in range(T):
            buf.obs[t].copy_(obs)
            if value_fn is not None:
                buf.value[t].copy_(value_fn(obs))
            act = policy(obs)
            buf.action[t].copy_(act)
            obs, rew, done = env.step(act)
            buf.reward[t].copy_(rew)
            buf.done[t].copy_(done)
            if value_fn is not None:
                buf.next_value[t].copy_(value_fn(obs))
            env.reset_done(write_obs=True)   # fresh observation for reset envs, step-time observation elsewhere
            obs = env.obs
        return buf
    own = dict(obs=env.obs, reward=env.reward, done=env.done)
    buf.obs[0].copy_(env.obs)
    try:
        for t in range(T):
            obs = buf.obs[t]
            if value_fn is not None:
                buf.value[t].copy_(value_fn(obs))
            buf.action[t].copy_(policy(obs))
            env.bind(reward=buf.reward[t], done=buf.done[t], obs=buf.obs[t + 1] if t + 1 < T else own["obs"])
            env.step(buf.action[t])        # copied into the env's own action buffer, which the kernel clamps
            if value_fn is not None:
                buf.next_value[t].copy_(value_fn(env.obs))
            env.reset_done(write_obs=True)
    finally:
        own["reward"].copy_(env.reward)
        own["done"].copy_(env.done)
        env.bind(**own)
    return buf


def compute_gae(buf: RolloutBuffer, gamma: float = 0.99, lmbda: float = 0.9):
    """Reverse scan over T on the GPU (sgb_gae); output lands in this rank's slot of the gather buffers."""
    if buf.reward.device.type != "cuda":
        raise _lib.SgbError("compute_gae runs only on CUDA tensors (no CPU fallback)")
    L = _lib.load_library()
    st = C.c_void_p(torch.cuda.current_stream(buf.reward.device).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _lib.check(L.sgb_gae(buf.T, buf.B, buf.N, p(buf.reward), p(buf.value), p(buf.next_value), p(buf.done),
                         gamma, lmbda, p(buf.advantage), p(buf.value_target), st), "sgb_gae")
    return buf.advantage, buf.value_target


def all_gather_advantages(buf: RolloutBuffer, group=None):
    """The single collective: in-place all-gather of [T, B_local, N] advantage and value-target buffers."""
    if buf.world == 1:
        return buf.adv_all, buf.target_all
    for full in (buf.adv_all, buf.target_all):
        dist.all_gather_into_tensor(full.view(-1), full[buf.rank].reshape(-1), group=group)
    return buf.adv_all, buf.target_all


def gae_allgather(buf: RolloutBuffer, gamma: float = 0.99, lmbda: float = 0.9, multicast: Optional[bool] = None):
    """GAE and the all-gather of its output as ONE kernel over peer memory (``sgb_gae_allgather``): replaces
    ``compute_gae`` + ``all_gather_advantages`` when the buffer was built with ``symmetric=True``.  `multicast=True`: store
    once to the NVSwitch multicast mapping instead of once per peer (an all-gather is bound by what every GPU RECEIVES,
    which multicast does not reduce; with 4-byte stores it measured slower than peer stores — 0.88 vs 0.48 ms on two
    B200s — so peer stores are the default).  Barriers on the
    current stream before (all ranks are done reading the last rollout's values) and after (all stores have landed)."""
    if buf.reward.device.type != "cuda":
        raise _lib.SgbError("gae_allgather runs only on CUDA tensors (no CPU fallback)")
    L = _lib.load_library()
    st = C.c_void_p(torch.cuda.current_stream(buf.reward.device).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    W = buf.world
    half = buf.adv_all.numel() * 4                      # bytes between the advantage and the value-target buffers
    if buf.symm is None:
        if W != 1:
            raise _lib.SgbError("gae_allgather over several ranks needs RolloutBuffer(symmetric=True)")
        bases, mc = [buf.adv_all.data_ptr()], 0
    else:
        bases = [int(x) for x in buf.symm.buffer_ptrs]
        mc = int(buf.symm.multicast_ptr or 0)
        if multicast and not mc:
            raise _lib.SgbError("this symmetric-memory handle has no multicast mapping")
        if not multicast:
            mc = 0
    adv = (C.c_void_p * W)(*bases)
    tgt = (C.c_void_p * W)(*[b + half for b in bases]) if buf.symm is not None else (C.c_void_p * W)(buf.target_all.data_ptr())
    if buf.symm is not None:
        buf.symm.barrier(channel=0)
    _lib.check(L.sgb_gae_allgather(buf.T, buf.B, buf.N, p(buf.reward), p(buf.value), p(buf.next_value), p(buf.done),
                                   gamma, lmbda, W, buf.rank, adv, tgt, C.c_void_p(mc or None),
                                   C.c_void_p((mc + half) if mc else None), st), "sgb_gae_allgather")
    if buf.symm is not None:
        buf.symm.barrier(channel=1)
    return buf.adv_all, buf.target_all


# ---- evaluation recordings ("out_td") --------------------------------------------------------------------------
# The reference's evaluation keeps the rollout's TensorDict ("out_td") and reads a handful of info entries from it
# (helper_common.py:581-611 trim_td, helper_training.py:1638-1698 reduce_out_td): leaves named
# ("agents", "info", key) with shape [num_envs, T, n_agents, F].  tensordict is not a dependency of this library, so
# the same layout is kept as a nested dict of tensors (``TensorDict(nested, batch_size=[B, T])`` wraps it as is).
OUT_TD_KEYS = ("pos", "rot", "vel", "ref", "ref_lanelet_ids", "is_collision_with_agents",
               "is_collision_with_lanelets")          # helper_common.py:591-599


def record_out_td(env, policy: Callable, T: int, keys=OUT_TD_KEYS, reset_done: bool = True):
    """Drive a VMAS(-like) environment of ``ScenarioRoadTrafficB200`` for T steps and stack what evaluation reads.

    `env`: ``vmas.Environment`` or ``VmasLikeEnvironment``; `policy(list of obs [B,D]) -> list of actions [B,2]``.
    Returns ``{"agents": {"observation": [B,T,N,D], "action": [B,T,N,2], "reward": [B,T,N,1],
    "info": {key: [B,T,N,F]}}, "done": [B,T,1]}`` — step-time values, i.e. what ``env.step`` returned at step t
    (TorchRL's ("next", ...) entries); bool entries stay bool (reduce_out_td's ``> 0.5`` accepts both).  Done envs
    are reset between steps like TorchRL's collector does (``reset_at``) unless `reset_done` is False."""
    agents = env.agents
    obs = [env.scenario.observation(a).clone() for a in agents]
    rec = dict(observation=[], action=[], reward=[], done=[], info={k: [] for k in keys})
    for _ in range(T):
        acts = policy(obs)
        rec["observation"].append(torch.stack(obs, dim=1))
        rec["action"].append(torch.stack([a.reshape(obs[0].shape[0], -1) for a in acts], dim=1))
        obs, rews, dones, infos = env.step(acts)
        rec["reward"].append(torch.stack(rews, dim=1).unsqueeze(-1))
        rec["done"].append(dones.reshape(-1, 1).clone())
        B = dones.shape[0]
        for k in keys:
            rec["info"][k].append(torch.stack([torch.as_tensor(i[k]).reshape(B, -1) for i in infos], dim=1))
        if reset_done and bool(dones.any()):
            for b in torch.nonzero(dones).flatten().tolist():
                env.reset_at(b)
            obs = [env.scenario.observation(a).clone() for a in agents]
    st = lambda xs: torch.stack(xs, dim=1)  # noqa: E731   [B, T, ...]
    return {"agents": {"observation": st(rec["observation"]), "action": st(rec["action"]), "reward": st(rec["reward"]),
                       "info": {k: st(v) for k, v in rec["info"].items()}},
            "done": st(rec["done"])}


def trim_out_td(out_td: dict, keys=OUT_TD_KEYS) -> dict:
    """helper_common.py:581-611 trim_td: keep only the selected ("agents", "info", key) leaves."""
    return {"agents": {"info": {k: out_td["agents"]["info"][k] for k in keys}}}


def reduce_out_td(out_td: dict, convert_collisions_to_bool: bool = True) -> dict:
    """helper_training.py:1638-1698 reduce_out_td for a single-env recording: (1, T, A, F) -> (T, A, F) leaves
    ``pos, rot (T, A), vel, is_collision_with_agents, is_collision_with_lanelets``; same errors on other shapes."""
    info = out_td["agents"]["info"]

    def squeeze_b1(name):
        if name not in info:
            raise KeyError(f"Missing key ('agents', 'info', '{name}') in out_td")
        x = info[name]
        if x.dim() < 4:
            raise ValueError(f"Expected a tensor with 4 dims (1,T,A,F), got shape {tuple(x.shape)}")
        if x.shape[0] != 1:
            raise ValueError(f"Expected leading batch size 1, got {x.shape[0]}")
        return x.squeeze(0)

    col_a, col_l = squeeze_b1("is_collision_with_agents"), squeeze_b1("is_collision_with_lanelets")
    if convert_collisions_to_bool:
        col_a, col_l = col_a > 0.5 if col_a.dtype != torch.bool else col_a, col_l > 0.5 if col_l.dtype != torch.bool else col_l
    return {"pos": squeeze_b1("pos"), "rot": squeeze_b1("rot").squeeze(-1), "vel": squeeze_b1("vel"),
            "is_collision_with_agents": col_a, "is_collision_with_lanelets": col_l}
