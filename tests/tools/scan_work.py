#!/usr/bin/env python
"""CPU tool (test infrastructure; uses the oracle to generate realistic states): work per agent-step of the kernels' own
polyline scans, counted by their host build (sgb_debug_scan_batch / sgb_debug_scan_counters, pruned mode) on states of
an oracle rollout with the bench's action distribution — segment evaluations, chunk boxes tested in the votes, exact
crossing predicates.  For sizing changes to the pruning logic without a GPU.

    python tests/tools/scan_work.py [scenario_type] [n_agents]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, ctypes as C
from oracle import oracle as O
from sigmarl_b200.lib import load_test_library as load_library
from sigmarl_b200.maps import MapLibrary
L=load_library()
st = sys.argv[1] if len(sys.argv) > 1 else "cpm_entire"
B, N = 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 8
m=MapLibrary(st); d=m.desc()
w=O.OracleWorld(st,B,N,mode="params",rew_method="distance")
for b in range(B): w.reset_env(b)
for b in range(B): w.refresh(b)
rng=np.random.default_rng(0)
ur=np.float32([1.0,31*np.pi/180])
cnt=np.zeros(8,np.int64); tot=0
L.sgb_debug_scan_counters(None,1)
for t in range(25):
    hint=w.idx_ref.copy().astype(np.int32).reshape(-1)           # carried closest index of the pre-step pose
    act=((rng.random((B,N,2),np.float32)*2-1)*ur).astype(np.float32)
    obs,rew,done,_=w.step(act,n_threads=8)
    path=w.path_id.copy().astype(np.int32).reshape(-1); pos=w.pos.copy().reshape(-1,2); psi=w.rot.copy().astype(np.float32).reshape(-1)
    xs=np.ascontiguousarray(pos[:,0]); ys=np.ascontiguousarray(pos[:,1]); out=np.zeros((B*N,16),np.float32)
    rc=L.sgb_debug_scan_batch(C.byref(d),B*N,path.ctypes.data,xs.ctypes.data,ys.ctypes.data,psi.ctypes.data,hint.ctypes.data,C.c_float(0.11),C.c_float(0.0535),2,out.ctypes.data)
    assert rc==0
    assert np.array_equal(out[:,1].astype(np.int32), w.idx_ref.reshape(-1)), "idx mismatch vs oracle"
    tot+=B*N
    for b in np.where(done)[0]:
        w.reset_env(int(b)); w.refresh(int(b))
c=np.zeros(8,np.int64); L.sgb_debug_scan_counters(c.ctypes.data,0)
print("agent-steps",tot)
print("per agent-step: centre seg evals %.1f | boundary seg evals %.1f (x5 points) | boxes tested %.1f | exact predicates %.2f | scans c/b %.1f/%.1f"%(c[0]/tot,c[1]/tot,c[2]/tot,c[3]/tot,c[4]/tot,c[5]/tot))
