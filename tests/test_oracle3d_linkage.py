"""SURVEY §8f row 4, second component (ORACLE ONLY so far, no CUDA kernel yet): moped3d's clustering stage
CLUSTER_LINKAGE_CPU. The C restatement (oracle/moped_linkage_oracle.c) against the class compiled unmodified from /root/reference
(oracle/ref3d_harness.cpp): clusters (which matches, in which order) identical to the strict-IEEE build in every case, and
identical to the build with the reference's own -ffast-math flags on well-separated data (merge decisions are comparisons of
similarities; only near-ties can flip under fast-math)."""
import numpy as np
import pytest

from oracle import oracle, ref3d

pytestmark = pytest.mark.skipif(not ref3d.available(), reason="oracle/_ref/libmoped3d_ref*.so not built (needs /root/reference at build time)")

W, H = 320, 240
K = np.array([262.0, 262.0, 160.0, 120.0])


def make_scene(seed, n_per=(30, 22), n_out=10, hallucinated=0.2):
    """Matches of ONE model seen twice (two instances at different places and depths) plus outliers; a depth map with the two
    instances as fronto-parallel patches over a slanted background, and its fill-distance map (0 = measured depth)."""
    rng = np.random.default_rng(seed)
    depth = (1.6 + 0.002 * np.arange(W)[None, :] + 0.001 * np.arange(H)[:, None]).astype(np.float32)
    dist = np.zeros((H, W), np.float32)
    xy, xyz, world = [], [], []
    for k, n in enumerate(n_per):
        cx, cy = (90 + 140 * k + rng.uniform(-10, 10), 110 + rng.uniform(-20, 20))
        z = 0.8 + 0.35 * k
        half = 38
        depth[int(cy) - half:int(cy) + half, int(cx) - half:int(cx) + half] = z
        pts = rng.uniform(-0.07, 0.07, (n, 3)).astype(np.float32)                    # model coordinates
        u = cx + pts[:, 0] * K[0] / z + rng.normal(0, 0.3, n)
        v = cy + pts[:, 1] * K[1] / z + rng.normal(0, 0.3, n)
        zz = z + pts[:, 2] * 0.1
        xy.append(np.stack([u, v], 1)); xyz.append(pts)
        world.append(np.stack([(u - K[2]) / K[0] * zz, (v - K[3]) / K[1] * zz, zz], 1))
    ou = rng.uniform([5, 5], [W - 5, H - 5], (n_out, 2))
    xy.append(ou); xyz.append(rng.uniform(-0.07, 0.07, (n_out, 3)))
    oz = depth[ou[:, 1].astype(int), ou[:, 0].astype(int)]
    world.append(np.stack([(ou[:, 0] - K[2]) / K[0] * oz, (ou[:, 1] - K[3]) / K[1] * oz, oz], 1))
    xy, xyz, world = (np.concatenate(a).astype(np.float32) for a in (xy, xyz, world))
    holes = rng.random((H, W)) < hallucinated
    dist[holes] = rng.uniform(1, 40, holes.sum()).astype(np.float32)
    perm = rng.permutation(len(xy))                                                  # matches arrive in query order, not by instance
    return xy[perm], xyz[perm], world[perm], depth, dist, perm


def as_sets(off, mem):
    return [tuple(mem[off[c]:off[c + 1]]) for c in range(len(off) - 1)]


@pytest.mark.parametrize("linkage", [1, 0, 2])
@pytest.mark.parametrize("use3d", [2, 1, 0])
def test_bit_identical_clusters_against_the_strict_build(linkage, use3d):
    for seed in range(4):
        xy, xyz, world, depth, dist, _ = make_scene(seed)
        ref3d.use_strict(True)
        try:
            ro, rm = ref3d.cluster_linkage(xy, xyz, world, depth, dist, use3d_filter=use3d, linkage_type=linkage)
        finally:
            ref3d.use_strict(False)
        oo, om = oracle.cluster_linkage(xy, xyz, world, depth, dist, use3d_filter=use3d, linkage_type=linkage)
        assert np.array_equal(ro, oo) and np.array_equal(rm, om), (seed, as_sets(ro, rm), as_sets(oo, om))


def test_default_parameters_recover_the_two_instances_like_the_reference():
    """moped3d/libmoped/src/config.hpp:45: CLUSTER_LINKAGE_CPU(0.1, 7, 2, 1, 0.0, 1, -1, -1), the reference's own flags."""
    agree = 0
    for seed in range(6):
        xy, xyz, world, depth, dist, perm = make_scene(10 + seed)
        ro, rm = ref3d.cluster_linkage(xy, xyz, world, depth, dist)
        oo, om = oracle.cluster_linkage(xy, xyz, world, depth, dist)
        agree += np.array_equal(ro, oo) and np.array_equal(rm, om)
        assert sorted(map(sorted, as_sets(ro, rm))) == sorted(map(sorted, as_sets(oo, om))), seed      # same partition at least
        # both instances come out as clusters made (almost) only of their own matches
        inst = np.where(perm < 30, 0, np.where(perm < 52, 1, 2))
        big = [c for c in as_sets(oo, om) if len(c) >= 15]
        assert len(big) >= 2
        for c in big[:2]:
            lab = np.bincount(inst[list(c)], minlength=3)
            assert lab.max() >= 0.85 * len(c)
    assert agree >= 5


def test_explicit_sigmas_small_and_degenerate_inputs():
    xy, xyz, world, depth, dist, _ = make_scene(3, n_per=(9, 0), n_out=0)
    for kw in (dict(sigma2d=20.0, sigma3d=0.05), dict(sigma2d=20.0), dict(cutoff=0.9, min_pts=2)):
        ref3d.use_strict(True)
        try:
            ro, rm = ref3d.cluster_linkage(xy, xyz, world, depth, dist, **kw)
        finally:
            ref3d.use_strict(False)
        oo, om = oracle.cluster_linkage(xy, xyz, world, depth, dist, **kw)
        assert np.array_equal(ro, oo) and np.array_equal(rm, om), kw
    # a single match, and features outside the image (coordinates are saturated for the depth path)
    oo, om = oracle.cluster_linkage(xy[:1], xyz[:1], world[:1], depth, dist, min_pts=0)
    assert len(oo) == 2 and list(om) == [0]


@pytest.mark.parametrize("linkage", [1, 0, 2])
def test_agglomeration_on_matrices_full_of_exact_ties(linkage):
    """hierarchicalCluster alone, reference vs restatement, on quantised symmetric matrices: every maximum is tied many times, so
    the scan order, the strict >, the stale entry of the merged-away cluster and the skipped list element all decide the result."""
    rng = np.random.default_rng(100 + linkage)
    for case in range(40):
        n = int(rng.integers(2, 48))
        levels = int(rng.integers(2, 6))
        K = rng.integers(0, levels + 1, (n, n)).astype(np.float32) / levels
        K = np.maximum(K, K.T)
        np.fill_diagonal(K, 1.0)
        cutoff = float(rng.choice([0.2, 0.5, 0.75, 1.0]))
        minpts = int(rng.integers(0, 4))
        for strict in (True, False):                     # comparisons and small-integer averages: both builds must agree exactly
            ref3d.use_strict(strict)
            try:
                ro, rm = ref3d.linkage_agglomerate(K, cutoff, minpts, linkage)
            finally:
                ref3d.use_strict(False)
            oo, om = oracle.linkage_agglomerate(K, cutoff, minpts, linkage)
            if linkage != 1 or strict:
                assert np.array_equal(ro, oo) and np.array_equal(rm, om), (case, n, cutoff, minpts, strict, as_sets(ro, rm), as_sets(oo, om))
