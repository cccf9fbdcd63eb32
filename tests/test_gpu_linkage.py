"""moped3d's clustering stage on a B200 (mc_cluster_linkage / mc_linkage_agglomerate) through the C ABI vs the
oracle (oracle/moped_linkage_oracle.c, bit-identical to the reference's strict build: tests/test_oracle3d_linkage.py).

Bars written here:
  * agglomeration on the ORACLE's own similarity matrix: identical clusters, identical member order, for every linkage type
    (the stage only compares similarities — no floating-point freedom);
  * similarity matrix: within 2e-6 absolute of the oracle's (CUDA vs glibc expf/atan2f);
  * end to end on the scenes of the oracle test: the same partition of the matches (identical arrays on at least 3 of 4
    scenes per setting: a near-tie between two similarities may be ordered differently by an ulp of expf).
"""
import numpy as np
import pytest

from oracle import oracle
from test_oracle3d_linkage import as_sets, make_scene

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("linkage", [1, 0, 2])
def test_agglomeration_bit_exact_on_the_oracles_matrix(gpu_ctx, linkage):
    for seed in range(4):
        xy, xyz, world, depth, dist, _ = make_scene(seed)
        K = oracle.linkage_similarity(xy, xyz, world, depth, dist)
        oo, om = oracle.linkage_agglomerate(K, 0.1, 7, linkage)
        go, gm = gpu_ctx.linkage_agglomerate(K, 0.1, 7, linkage)
        assert np.array_equal(oo, go) and np.array_equal(om, gm), (seed, as_sets(oo, om), as_sets(go, gm))
    # other cutoffs / MinPts, and a matrix with exact ties everywhere (first pair in scan order must win)
    rng = np.random.default_rng(1)
    K = rng.integers(0, 4, (40, 40)).astype(np.float32) / 4
    K = np.maximum(K, K.T); np.fill_diagonal(K, 1.0)
    for cutoff, minpts in ((0.5, 0), (0.75, 2), (0.2, 5)):
        oo, om = oracle.linkage_agglomerate(K, cutoff, minpts, linkage)
        go, gm = gpu_ctx.linkage_agglomerate(K, cutoff, minpts, linkage)
        assert np.array_equal(oo, go) and np.array_equal(om, gm), (cutoff, minpts)


@pytest.mark.parametrize("use3d", [2, 1, 0])
def test_similarity_matrix_and_clusters_end_to_end(gpu_ctx, use3d):
    same = 0
    for seed in range(4):
        xy, xyz, world, depth, dist, _ = make_scene(seed)
        K = oracle.linkage_similarity(xy, xyz, world, depth, dist, use3d_filter=use3d)
        go, gm, GK = gpu_ctx.cluster_linkage(xy, xyz, world, depth, dist, use3d_filter=use3d, want_similarity=True)
        assert np.abs(GK - K).max() < 2e-6, np.abs(GK - K).max()
        assert np.array_equal(GK, GK.T)
        oo, om = oracle.cluster_linkage(xy, xyz, world, depth, dist, use3d_filter=use3d)
        assert sorted(map(sorted, as_sets(oo, om))) == sorted(map(sorted, as_sets(go, gm))), seed
        same += np.array_equal(oo, go) and np.array_equal(om, gm)
    assert same >= 3


def test_explicit_sigmas_and_small_inputs(gpu_ctx):
    xy, xyz, world, depth, dist, _ = make_scene(3, n_per=(9, 0), n_out=0)
    for kw in (dict(sigma2d=20.0, sigma3d=0.05), dict(cutoff=0.9, min_pts=2)):
        oo, om = oracle.cluster_linkage(xy, xyz, world, depth, dist, **kw)
        go, gm = gpu_ctx.cluster_linkage(xy, xyz, world, depth, dist, **kw)
        assert sorted(map(sorted, as_sets(oo, om))) == sorted(map(sorted, as_sets(go, gm))), kw
    go, gm = gpu_ctx.cluster_linkage(xy[:1], xyz[:1], world[:1], depth, dist, min_pts=0)
    assert list(go) == [0, 1] and list(gm) == [0]
