#!/usr/bin/env python
"""CPU tool: randomized hunt for a counterexample to "pruned scan == exhaustive scan" with the kernels' own scan source
(host build, sgb_debug_scan_batch) — 6 rounds x 18 maps x 20 000 poses of three kinds: near the centre lines with noise,
anywhere on the map with arbitrary headings and garbage hints, exactly on centre / boundary points with headings along
the path or its normal (collinear edges, exact ties).  Last run: 2 160 000 poses, 0 mismatches (30 s).

    python tests/tools/scan_campaign.py
"""
import numpy as np, ctypes as C, sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from sigmarl_b200.lib import load_test_library as load_library
from sigmarl_b200.maps import MapLibrary, available_scenarios
import test_abi_and_host as T
L=load_library()
rng=np.random.default_rng(12345)
t0=time.time(); total=0
for rep in range(6):
  for st in available_scenarios():
    m=MapLibrary(st)
    n=20000
    path,pos,psi,hint=T._scan_poses(m,n,rng)
    kind=rep%3
    if kind==1:   # anywhere on the map, any heading, garbage hints
        pos=rng.uniform([-0.3,-0.3],[m.world_x_dim+0.3,m.world_y_dim+0.3],(n,2)).astype(np.float32)
        psi=rng.uniform(-7,7,n).astype(np.float32); hint=rng.integers(-5,300,n).astype(np.int32)
    elif kind==2: # exactly on centre / boundary points, headings exactly along the path or its normal
        k=(rng.random(n)*(m.n_center[path]-1)).astype(np.int64)
        pos=m.center_xy[m.center_off[path]+k].astype(np.float32).copy()
        half=n//2
        pos[:half]=m.left_xy[np.minimum(m.left_off[path[:half]]+k[:half], m.left_off[path[:half]+1]-1)]
        yaw_off=np.concatenate([[0],np.cumsum(m.n_center-1)])
        yaw=m.center_yaw[np.minimum(yaw_off[path]+k, yaw_off[path+1]-1)]
        psi=(yaw+rng.integers(0,4,n)*np.float32(np.pi/2)).astype(np.float32)
    a=T._scan_batch(L,m,path,pos,psi,hint,0); b=T._scan_batch(L,m,path,pos,psi,hint,1)
    total+=n
    if not np.array_equal(a.view(np.uint32),b.view(np.uint32)):
        bad=np.where((a.view(np.uint32)!=b.view(np.uint32)).any(1))[0]
        print("MISMATCH",st,"kind",kind,len(bad),"first",bad[0],path[bad[0]],pos[bad[0]],psi[bad[0]],hint[bad[0]]); print(a[bad[0]]); print(b[bad[0]])
print("done",total,"poses in %.0f s"%(time.time()-t0))
