#!/usr/bin/env python
"""TEST INFRASTRUCTURE — record what the UNMODIFIED reference draws in full environment resets.

`reset_world_at(env)` places the agents one after the other: path ~ U{paths of the env's set}, point ~ U[3, n/2), speed ~
U(0, v_max), re-drawn until the agent keeps `reset_agent_min_distance` from the agents placed before it
(world_state_rt_sim.py:215-311); on cpm_mixed the env first draws its path set ~ multinomial(cpm_scenario_probabilities)
(:313-358).  The device reset of the library is distribution-equivalent, not stream-equivalent, to this; the fixtures
written here (tests/golden/resets/*.npz) are the reference sample its draws are compared with
(tests/test_gpu_parity.py::test_device_reset_draws_follow_the_references_distribution).

Re-run (here only):  python oracle/gen_reset_draws.py
"""
import os
import sys

os.environ["CICD_TESTING"] = "true"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
import install_shims  # noqa: E402,F401

import numpy as np  # noqa: E402
import torch  # noqa: E402
from vmas.simulator.environment import Environment  # noqa: E402
from sigmarl.helper_common import Parameters  # noqa: E402
from sigmarl.scenarios.road_traffic import ScenarioRoadTraffic  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "resets")
CONFIGS = {
    "cpm_entire_N4": dict(st="cpm_entire", N=4, B=8, rounds=300, seed=101),
    # (merge-out is left out: 2 of its 31 spawn points have no feasible partner point, so the reference's unbounded
    # rejection loop spins forever in ~6 % of the resets of a two-agent env there)
    "cpm_mixed_N2_sets": dict(st="cpm_mixed", N=2, B=8, rounds=200, seed=102, extra=dict(cpm_scenario_probabilities=[0.4, 0.6, 0.0])),
    "on_ramp_2_N6": dict(st="on_ramp_2_multilane", N=6, B=8, rounds=150, seed=103),
}


def run(name, st, N, B, rounds, seed, extra=None):
    torch.manual_seed(seed)
    sc = ScenarioRoadTraffic()
    p = Parameters.from_json("/root/reference/sigmarl/config.json")
    p.scenario_type, p.n_agents, p.num_vmas_envs = st, N, B
    for k, v in (extra or {}).items():
        setattr(p, k, v)
    sc.parameters = p
    env = Environment(sc, num_envs=B, device="cpu", max_steps=128)
    ws, agents = sc.world_state, env.agents
    rec = {k: [] for k in ("path_id", "point_id", "scenario_id", "speed", "pos")}
    for r in range(rounds):
        for b in range(B):
            env.reset_at(b)
        rec["path_id"].append(ws.ref_paths_agent_related.path_id.clone().numpy())
        rec["point_id"].append(ws.ref_paths_agent_related.point_id.clone().numpy())
        rec["scenario_id"].append(ws.ref_paths_agent_related.scenario_id.clone().numpy())
        rec["speed"].append(torch.stack([a.state.speed.squeeze(-1) for a in agents], 1).clone().numpy())
        rec["pos"].append(torch.stack([a.state.pos for a in agents], 1).clone().numpy())
    out = {k: np.concatenate(v) for k, v in rec.items()}        # [rounds * B, N, ...]
    out["cfg_scenario_type"], out["cfg_N"] = np.asarray(st), np.asarray(N)
    out["cfg_probabilities"] = np.asarray((extra or {}).get("cpm_scenario_probabilities", [1.0, 0.0, 0.0]), np.float32)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if not k.startswith("cfg")})


if __name__ == "__main__":
    for name, c in CONFIGS.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        run(name, **c)
