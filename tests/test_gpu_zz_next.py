"""GPU parity for the features added after the round's last hardware session (fixtures in tests/golden/next/):
``reset_agent_fixed_duration`` (road_traffic.py:1388-1393) and the MTV agent distance (``is_use_mtv_distance``,
helper_scenario.py:1030-1138).  The oracle is pinned on the same fixtures on the CPU (test_oracle_golden.py); these
run the CUDA path through the C-ABI against them, with exactly the checks of test_gpu_parity.py.  The file sorts after
the rest of the suite on purpose: it is the part that has not been on a B200 yet.
"""
import os

import numpy as np
import pytest
import torch

from conftest import golden_files
import test_gpu_parity as P

pytestmark = pytest.mark.gpu

NEXT = golden_files("next")
_ids = lambda p: os.path.basename(p)[:-4]  # noqa: E731


@pytest.mark.parametrize("path", NEXT, ids=_ids)
@pytest.mark.parametrize("exhaustive", [False, True], ids=["pruned", "exhaustive"])
def test_cuda_matches_reference_goldens_next(path, exhaustive):
    P.test_cuda_matches_reference_goldens(path, exhaustive)


@pytest.mark.parametrize("path", NEXT, ids=_ids)
def test_cuda_reset_obs_matches_reference_next(path):
    P.test_cuda_reset_obs_matches_reference(path)


@pytest.mark.parametrize("path", NEXT, ids=_ids)
def test_facade_info_matches_reference_goldens_next(path):
    P.test_facade_info_matches_reference_goldens(path)


def test_fixed_duration_ends_envs_periodically_free_running():
    """dt 0.1, reset_agent_fixed_duration 1 s, testing mode (only the clock ends an env): every env is done exactly
    at timer.step 10, 20, ... and the device reset restarts the clock."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    B, N = 512, 4
    cfg = EnvConfig(scenario_type="cpm_entire", n_agents=N, mode="params", max_steps=64, is_testing_mode=True,
                    reset_agent_fixed_duration=1)
    env = RoadTrafficEnv(cfg, num_envs=B, device="cuda:0", seed=9)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(1)
    ur = torch.as_tensor(P.UR).cuda()
    for t in range(1, 35):
        act = (torch.rand(B, N, 2, generator=gen, device="cuda") * 2 - 1) * ur
        _, _, done = env.step(act)
        want = (t % 10 == 0)
        assert bool(done.bool().all()) == want and bool(done.bool().any()) == want, t
        env.reset_done()
        # a fixed-duration reset zeroes the clock (road_traffic.py:877), so the period restarts
        assert int(env.step_count.max()) == (0 if want else t % 10)
