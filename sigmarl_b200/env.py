"""RoadTrafficEnv — batched tensor front-end of the fused CUDA environment step.

Owns the device buffers (torch tensors; the C library only sees ``data_ptr()``) and one ``sgb_ctx``.
State layout (env-major, agent-minor — DESIGN.md "Data layout in HBM"):
    pose [B,N,4]  x, y, psi, v            aux [B,N,4]  delta, vx, vy, beta
    path_id [B,N] int32                   carry [B,N,4]  d_ref, min dL, min dR, idx_ref (pre-step pose)
    action [B,N,2]  obs [B,N,D]  reward [B,N]  done [B] u8  agent_flags [B,N] u8  step_count [B] i32
    info [B,N,16] (optional; layout in include/sigmarl_b200.h)  task_tries / task_success [B] i32 (optional)
"""
import ctypes as C

import numpy as np
import torch

from . import lib as _lib
from .config import EnvConfig
from .maps import MapLibrary


class RoadTrafficEnv:
    def __init__(self, config: EnvConfig = None, num_envs: int = 32, device="cuda:0", seed: int = 0,
                 env_offset: int = 0, debug: bool = False, max_reset_tries: int = 64, info: bool = False,
                 **cfg_kwargs):
        self.config = config or EnvConfig(**cfg_kwargs)
        self.L = _lib.load_library()                      # raises if the .so is missing
        if not torch.cuda.is_available():
            raise _lib.SgbError("RoadTrafficEnv needs a CUDA device: the environment step exists only as CUDA "
                                "kernels (no CPU fallback)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.SgbError(f"device must be a CUDA device, got {device}")
        self.map = MapLibrary(self.config.scenario_type)
        self.cfg = self.config.lower(self.map)
        r = self.config.resolved(self.config.lane_width(self.map), self.map.default_n_agents)
        self.B, self.N = int(num_envs), int(r["n_agents"])
        if not 1 <= self.N <= _lib.SGB_MAX_AGENTS:
            raise ValueError(f"n_agents must be in [1, {_lib.SGB_MAX_AGENTS}]")
        self.dt = r["dt"]
        self.seed, self.env_offset, self.epoch = int(seed), int(env_offset), 0
        self.max_reset_tries = int(max_reset_tries)
        # paths a reset draws from: one range, or (cpm_mixed with several weighted sets) a set drawn per env
        rng = self.map.default_path_range(self.config.cpm_scenario_probabilities)
        self.per_env_path_sets = rng is None
        self.path_lo, self.path_hi = (-1, 0) if rng is None else rng
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._ctx = C.c_void_p()
        d = self.map.desc()
        with torch.cuda.device(idx):
            _lib.check(self.L.sgb_create(C.byref(self._ctx), idx, C.byref(d), C.byref(self.cfg)), "sgb_create")
        if self.cfg.obs_flags & _lib.SGB_OBS_MASK_LANELETS:      # lanelet table for the lanelet-relation observation mask
            m = self.map
            with torch.cuda.device(idx):
                _lib.check(self.L.sgb_set_lanelets(self._ctx, len(m.lanelet_off) - 1, m.lanelet_xy.ctypes.data,
                                                   m.lanelet_off.ctypes.data, m.lanelet_adj.ctypes.data), "sgb_set_lanelets")
        _lib.check(self.L.sgb_set_env_offset(self._ctx, self.env_offset), "sgb_set_env_offset")
        if self.per_env_path_sets:
            lo, hi, pr = self.map.path_sets(self.config.cpm_scenario_probabilities)
            _lib.check(self.L.sgb_set_path_sets(self._ctx, len(lo), lo.ctypes.data, hi.ctypes.data, pr.ctypes.data),
                       "sgb_set_path_sets")
        self.D = self.L.sgb_obs_dim(self._ctx)
        B, N, dev = self.B, self.N, self.device
        z = lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype, device=dev)  # noqa: E731
        self.pose, self.aux, self.carry = z(B, N, 4), z(B, N, 4), z(B, N, 4)
        self.path_id = z(B, N, dtype=torch.int32)
        self.action = z(B, N, 2)
        self.step_count = z(B, dtype=torch.int32)
        self.obs, self.reward = z(B, N, self.D), z(B, N)
        self.done = z(B, dtype=torch.uint8)
        self.agent_flags = z(B, N, dtype=torch.uint8)
        self.collide_with = z(B, N, dtype=torch.int32)
        self.dbg = z(B, N, 16) if debug else None
        # info(agent) extras + evaluation counters (road_traffic.py:1489-1635, :998-1035): only when asked for
        self.info = z(B, N, _lib.SGB_INFO_DIM) if info else None
        self.task_tries = z(B, dtype=torch.int32) if info else None
        self.task_success = z(B, dtype=torch.int32) if info else None
        self.n_failed = z(1, dtype=torch.int32)
        # path set of every env (reference: ref_paths_agent_related.scenario_id - 1 on cpm_mixed, 0 elsewhere); written
        # by the device reset when it draws a set per env, otherwise constant
        self.scenario_id = z(B, dtype=torch.int32)
        if not self.per_env_path_sets:
            self.scenario_id.fill_(int(self.map.set_of_path(np.asarray([self.path_lo]))[0]))
        self.nan_flags = z(1, dtype=torch.int32)           # sticky health word (bit 0: a step produced NaN / inf)
        self._buf = _lib.Buffers()
        for name in _lib.BUFFER_FIELDS:
            t = getattr(self, name)
            setattr(self._buf, name, t.data_ptr() if t is not None else None)
        self._h = None  # pinned host staging for step_host

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.L.sgb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self.L.sgb_launch_count(self._ctx))

    @property
    def map_bytes(self):
        return int(self.L.sgb_map_bytes(self._ctx))

    def bind(self, **tensors):
        """Point kernel inputs / outputs at caller-owned tensors of the same shape and dtype (e.g. slices of a
        rollout buffer), so that a step writes them in place: ``env.bind(obs=buf.obs[t + 1], reward=buf.reward[t])``.
        Only ``obs``, ``reward``, ``done`` and ``action`` may be re-bound; state buffers stay owned by the env."""
        for name, t in tensors.items():
            if name not in ("obs", "reward", "done", "action"):
                raise ValueError(f"cannot re-bind {name!r}")
            cur = getattr(self, name)
            if t.shape != cur.shape or t.dtype != cur.dtype or t.device != cur.device or not t.is_contiguous():
                raise ValueError(f"bind({name}): need a contiguous {tuple(cur.shape)} {cur.dtype} tensor on {cur.device}")
            setattr(self, name, t)
            setattr(self._buf, name, t.data_ptr())

    # ------------------------------------------------------------------ the path
    def step(self, action: torch.Tensor = None, auto_reset: bool = False):
        """One fused-kernel environment step.  Returns views (obs [B,N,D], reward [B,N], done [B] uint8).
        auto_reset: done envs are reset right away and `obs` then holds their post-reset observation (what the policy
        acts on next), reward / done stay the step's."""
        if action is not None:
            self.action.copy_(action.reshape(self.B, self.N, 2))
        _lib.check(self.L.sgb_step(self._ctx, self.B, self.N, C.byref(self._buf), self._stream()), "sgb_step")
        if auto_reset:
            self.reset_done(write_obs=True)
        return self.obs, self.reward, self.done

    def reset_done(self, write_obs: bool = True):
        """Masked device-side reset of done envs + respawn of agents that crossed an entry/exit segment."""
        self.epoch += 1
        _lib.check(self.L.sgb_reset(self._ctx, self.B, self.N, C.byref(self._buf), self.path_lo, self.path_hi,
                                    self.seed, self.epoch, self.env_offset, self.max_reset_tries, int(write_obs),
                                    C.c_void_p(self.n_failed.data_ptr()), self._stream()), "sgb_reset")

    def reset_masked(self, env_mask: torch.Tensor = None, agent_mask: torch.Tensor = None, write_obs: bool = True,
                     path_range=None):
        """Explicit selection: fully reset the envs of `env_mask` [B], respawn the agents of `agent_mask` [B,N]
        (whatever the map / mode) — ``reset_world_at(env_index, agent_index)`` called from outside the step.
        `path_range` = (lo, hi) restricts the paths drawn from for this call (``predefined_ref_path_idx`` respawns,
        world_state_rt_sim.py:241-242); default: the env's own range."""
        self.epoch += 1
        path_lo, path_hi = (self.path_lo, self.path_hi) if path_range is None else (int(path_range[0]), int(path_range[1]))
        if path_range is not None and not (0 <= path_lo < path_hi <= self.map.n_paths):
            raise ValueError(f"path_range {path_range} outside [0, {self.map.n_paths}]")
        prep = lambda m: None if m is None else m.to(device=self.device, dtype=torch.uint8).contiguous()  # noqa: E731
        em, am = prep(env_mask), prep(agent_mask)
        ptr = lambda m: C.c_void_p(m.data_ptr()) if m is not None else None  # noqa: E731
        _lib.check(self.L.sgb_reset_masked(self._ctx, self.B, self.N, C.byref(self._buf), ptr(em), ptr(am), path_lo,
                                           path_hi, self.seed, self.epoch, self.env_offset, self.max_reset_tries,
                                           int(write_obs), C.c_void_p(self.n_failed.data_ptr()), self._stream()),
                   "sgb_reset_masked")

    def reset(self):
        """Environment.reset(): (re)place every agent of every env; returns the fresh observation."""
        self.epoch += 1
        _lib.check(self.L.sgb_reset_all(self._ctx, self.B, self.N, C.byref(self._buf), self.path_lo, self.path_hi,
                                        self.seed, self.epoch, self.env_offset, self.max_reset_tries,
                                        C.c_void_p(self.n_failed.data_ptr()), self._stream()), "sgb_reset_all")
        self.done.zero_()
        return self.obs

    def refresh(self, env_mask: torch.Tensor = None, write_obs: bool = False):
        m = None
        if env_mask is not None:
            m = env_mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _lib.check(self.L.sgb_refresh(self._ctx, self.B, self.N, C.byref(self._buf),
                                      C.c_void_p(m.data_ptr()) if m is not None else None, int(write_obs),
                                      self._stream()), "sgb_refresh")
        return self.obs

    def place(self, path, point, speed, agent_mask=None):
        """Put agents at (path, point) with `speed` (parity mode: the caller supplies the reset draws)."""
        dev = self.device
        path = torch.as_tensor(np.asarray(path), dtype=torch.int32, device=dev).contiguous()
        point = torch.as_tensor(np.asarray(point), dtype=torch.int32, device=dev).contiguous()
        speed = torch.as_tensor(np.asarray(speed), dtype=torch.float32, device=dev).contiguous()
        m = None
        if agent_mask is not None:
            m = torch.as_tensor(np.asarray(agent_mask), device=dev).to(torch.uint8).contiguous()
        _lib.check(self.L.sgb_place(self._ctx, self.B, self.N, C.byref(self._buf),
                                    C.c_void_p(m.data_ptr()) if m is not None else None,
                                    C.c_void_p(path.data_ptr()), C.c_void_p(point.data_ptr()),
                                    C.c_void_p(speed.data_ptr()), self._stream()), "sgb_place")

    def set_state(self, pos, rot, speed, steering, path_id, step_count=None, write_obs=False):
        """Teacher forcing: inject a state, then rebuild everything derived from it (sgb_refresh)."""
        dev = self.device
        t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev)  # noqa: E731
        self.pose[..., 0:2] = t(pos)
        self.pose[..., 2] = t(rot)
        self.pose[..., 3] = t(speed)
        self.aux[..., 0] = t(steering)
        self.path_id.copy_(torch.as_tensor(np.asarray(path_id), dtype=torch.int32, device=dev))
        if step_count is not None:
            self.step_count.copy_(torch.as_tensor(np.asarray(step_count), dtype=torch.int32, device=dev))
        return self.refresh(write_obs=write_obs)

    def set_pose_history(self, from_reset):
        """Teacher forcing with ``is_observe_distance_to_boundaries=False``: say for which agents [B,N] the injected
        pose was written by a reset / respawn (True) rather than by a step (False).  The reference samples the nearing
        boundary points differently in the two cases (world_state_rt.py:531-576 vs :686-725) and agents >= 1 observe
        the previous write; the library keeps that one bit in ``carry.w`` (bit 30).  ``set_state`` / ``refresh``
        mark every agent as reset, a step clears the mark."""
        m = torch.as_tensor(np.asarray(from_reset), device=self.device).to(torch.bool).reshape(self.B, self.N)
        w = self.carry.view(torch.int32)[..., 3]
        w.copy_(torch.where(m, w | _lib.CARRY_FRESH_BIT, w & _lib.CARRY_IDX_MASK))

    @property
    def pose_from_reset(self):
        """[B,N] bool: the mark described in ``set_pose_history`` (always False for the other layouts)."""
        return (self.carry.view(torch.int32)[..., 3] & _lib.CARRY_FRESH_BIT) != 0

    def step_host(self, h_action: torch.Tensor, reset_done: bool = False):
        """End-to-end step with HOST buffers through sgb_step_host (H2D action, D2H obs/reward/done inside).
        reset_done: sgb_step_reset_host — done envs are reset inside the same pipeline and the returned observation
        is the one the policy acts on next (post-reset for the envs that finished, step-time for the others)."""
        if self._h is None:
            pin = lambda *s, dtype=torch.float32: torch.empty(*s, dtype=dtype).pin_memory()  # noqa: E731
            self._h = dict(obs=pin(self.B, self.N, self.D), reward=pin(self.B, self.N),
                           done=pin(self.B, dtype=torch.uint8))
        h = self._h
        assert h_action.device.type == "cpu" and h_action.dtype == torch.float32 and h_action.is_contiguous()
        ptrs = (C.c_void_p(h_action.data_ptr()), C.c_void_p(h["obs"].data_ptr()), C.c_void_p(h["reward"].data_ptr()),
                C.c_void_p(h["done"].data_ptr()))
        if reset_done:
            self.epoch += 1
            _lib.check(self.L.sgb_step_reset_host(self._ctx, self.B, self.N, C.byref(self._buf), *ptrs, self.path_lo,
                                                  self.path_hi, self.seed, self.epoch, self.env_offset,
                                                  self.max_reset_tries, C.c_void_p(self.n_failed.data_ptr()),
                                                  self._stream()), "sgb_step_reset_host")
        else:
            _lib.check(self.L.sgb_step_host(self._ctx, self.B, self.N, C.byref(self._buf), *ptrs, self._stream()),
                       "sgb_step_host")
        return h["obs"], h["reward"], h["done"]

    # ------------------------------------------------------------------ reference-named views
    @property
    def pos(self): return self.pose[..., 0:2]
    @property
    def rot(self): return self.pose[..., 2]
    @property
    def speed(self): return self.pose[..., 3]
    @property
    def steering(self): return self.aux[..., 0]
    @property
    def vel(self): return self.aux[..., 1:3]
    @property
    def sideslip_angle(self): return self.aux[..., 3]
