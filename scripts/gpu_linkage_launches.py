import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from moped_b200 import capi
from test_oracle3d_linkage import make_scene
ctx = capi.Context(0)
xy, xyz, world, depth, dist, _ = make_scene(1, n_per=(120, 80), n_out=40)
for _ in range(2):
    ctx.cluster_linkage(xy, xyz, world, depth, dist)
