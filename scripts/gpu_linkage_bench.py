"""Times mc_cluster_linkage (host buffers in, clusters out) against the compiled reference on the host for a few match counts."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from moped_b200 import capi
from oracle import ref3d
from test_oracle3d_linkage import make_scene

ctx = capi.Context(0)
for n_per, n_out in (((30, 22), 10), ((120, 80), 40), ((300, 200), 100)):
    xy, xyz, world, depth, dist, _ = make_scene(1, n_per=n_per, n_out=n_out)
    n = len(xy)
    for _ in range(3):
        out = ctx.cluster_linkage(xy, xyz, world, depth, dist)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        out = ctx.cluster_linkage(xy, xyz, world, depth, dist)
    gpu_ms = (time.perf_counter() - t0) / reps * 1e3
    cpu_ms = None
    if ref3d.available():
        t0 = time.perf_counter()
        r = ref3d.cluster_linkage(xy, xyz, world, depth, dist)
        cpu_ms = (time.perf_counter() - t0) * 1e3
    ctx.set_option("linkage_cached", 1)
    for _ in range(3):
        out_c = ctx.cluster_linkage(xy, xyz, world, depth, dist)
    t0 = time.perf_counter()
    for _ in range(reps):
        out_c = ctx.cluster_linkage(xy, xyz, world, depth, dist)
    cached_ms = (time.perf_counter() - t0) / reps * 1e3
    ctx.set_option("linkage_cached", 0)
    assert np.array_equal(out[0], out_c[0]) and np.array_equal(out[1], out_c[1]), "cached agglomeration changed the clusters"
    print(json.dumps(dict(n_matches=n, gpu_ms_per_call=gpu_ms, gpu_ms_per_call_cached_maxima=cached_ms, reference_cpu_ms_per_call=cpu_ms, clusters=len(out[0]) - 1,
                          note="mc_cluster_linkage with host buffers (two 320x240 maps + matches up, clusters down) vs CLUSTER_LINKAGE_CPU, one model")))
