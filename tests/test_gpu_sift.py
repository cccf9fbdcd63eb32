"""Feature extraction on a B200 (mc_sift_extract) through the C ABI vs the oracle
(oracle/moped_sift_oracle.c, pinned to the compiled reference by tests/test_sift_oracle.py) and vs the golden vectors
the compiled reference produced.

Bars written here:
  * scale-space (Gaussian and DoG planes): BIT-EXACT vs the oracle (same sums, no FMA contraction);
  * keypoints vs the oracle: same count, same order, coord2D bit-exact (it only depends on the DoG stack), scale to 1e-6
    relative (powf);
    orientation within 1e-4 rad, descriptor components within 2e-4 (libm vs CUDA expf/atan2f/sincos and the
    order-independent fixed-point bin sums); >= 99.5 % of the keypoints must satisfy this — a vote landing on the other
    side of a histogram-bin edge by one ulp of atan2f may flip a secondary orientation peak;
  * vs the reference's golden vectors: the tolerant gate of tests/sift_util.py (the reference is -ffast-math).
"""
import os

import numpy as np
import pytest

from sift_util import assert_same_keypoints, match_keypoints

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "sift_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def compare_with_oracle(got, want, min_frac=0.995):
    (xy, so, desc), (oxy, oso, odesc) = got, want
    n, m = len(xy), len(oxy)
    assert abs(n - m) <= max(1, 0.005 * m), (n, m)
    if n == m and np.array_equal(xy, oxy):
        assert np.allclose(so[:, 0], oso[:, 0], rtol=1e-6, atol=0)          # scale = 1.6 * powf(2, .): one ulp of powf
        dori = np.abs(so[:, 1] - oso[:, 1])
        ddesc = np.abs(desc - odesc).max(axis=1)
        ok = (dori < 1e-4) & (ddesc < 2e-4)
        assert ok.mean() >= min_frac, (ok.mean(), np.sort(dori)[-3:], np.sort(ddesc)[-3:])
        return True
    idx, dd = match_keypoints(xy, desc, oxy, odesc, tol_px=1e-6)          # a flipped orientation peak: compare as sets
    assert (idx >= 0).mean() >= min_frac and (dd[idx >= 0] < 2e-4).mean() >= min_frac
    return False


def test_scale_space_is_bit_exact(gpu_ctx, gold, oracle_mod):
    name = "bag0_crop_double"
    im = gold[f"{name}/image"]
    gpu_ctx.sift(im, True)
    n_oct = len(oracle_mod.sift_octave_dims(*im.shape, True))
    for octv in (0, 1, n_oct - 1):
        gauss, dog, _ = oracle_mod.sift_debug(im, True, octv)
        for i in range(6):
            assert np.array_equal(gpu_ctx.sift_plane(0, octv, 0, i), gauss[i]), (octv, "gauss", i)
        for i in range(5):
            assert np.array_equal(gpu_ctx.sift_plane(0, octv, 1, i), dog[i]), (octv, "dog", i)


def test_keypoints_match_oracle_on_every_golden_case(gpu_ctx, gold, oracle_mod):
    same = 0
    for name in gold["names"]:
        im, dbl = gold[f"{name}/image"], bool(gold[f"{name}/double"])
        got = gpu_ctx.sift(im, dbl)
        same += compare_with_oracle(got, oracle_mod.sift(im, dbl))
        assert np.allclose(np.linalg.norm(got[2], axis=1), 1.0, atol=1e-5)
    assert same >= 4


def test_keypoints_match_reference_golden_vectors(gpu_ctx, gold):
    for name in gold["names"]:
        xy, so, desc = gpu_ctx.sift(gold[f"{name}/image"], bool(gold[f"{name}/double"]))
        assert_same_keypoints(xy, desc, gold[f"{name}/xy"], gold[f"{name}/desc"])


def test_batch_equals_single_images_and_is_deterministic(gpu_ctx, gold):
    base = gold["bag4_full_double/image"]
    frames = np.stack([base, base[::-1].copy(), np.roll(base, 37, axis=1), base])
    singles = [gpu_ctx.sift(f, True) for f in frames]
    batch = gpu_ctx.sift(frames, True)
    again = gpu_ctx.sift(frames, True)
    for s, b, a in zip(singles, batch, again):
        for x, y, z in zip(s, b, a):
            assert np.array_equal(x, y) and np.array_equal(y, z)
    assert np.array_equal(batch[0][2], batch[3][2])
    assert len(batch[0][0]) > 300


def test_capacity_and_degenerate_inputs(gpu_ctx, gold):
    from moped_b200 import capi
    im = gold["bag0_crop_double/image"]
    with pytest.raises(capi.MopedCudaError, match="max_keypoints"):
        gpu_ctx.sift(im, True, max_keypoints=16)
    flat = np.full((64, 80), 128, np.uint8)
    assert len(gpu_ctx.sift(flat, True)[0]) == 0
    assert len(gpu_ctx.sift(np.zeros((12, 40), np.uint8), False)[0]) == 0       # no octave at all (rows <= 12)
    xy, so, desc = gpu_ctx.sift(im, True)                                        # the context still works afterwards
    assert len(xy) == len(gold["bag0_crop_double/xy"]) or abs(len(xy) - len(gold["bag0_crop_double/xy"])) <= 2


def test_fused_blur_equals_two_pass_kernels(gpu_ctx, gold):
    """The shared-memory fused Gaussian kernel and the separate row/column kernels sum the taps in the same order."""
    im = gold["ex2_odd_double/image"]
    try:
        fused = gpu_ctx.sift(im, True)
        planes = [gpu_ctx.sift_plane(0, o, 1, 4) for o in (0, 2)]
        gpu_ctx.set_option("sift_two_pass", 1)
        two = gpu_ctx.sift(im, True)
        planes2 = [gpu_ctx.sift_plane(0, o, 1, 4) for o in (0, 2)]
    finally:
        gpu_ctx.set_option("sift_two_pass", 0)
    for a, b in zip(fused, two):
        assert np.array_equal(a, b)
    for a, b in zip(planes, planes2):
        assert np.array_equal(a, b)


def test_descriptor_kernels_agree(gpu_ctx, gold):
    """Warp-per-keypoint scatter (default) and cell-gather descriptor kernels add the same terms in different orders."""
    im = gold["bag0_crop_double/image"]
    try:
        a = gpu_ctx.sift(im, True)
        gpu_ctx.set_option("sift_describe_gather", 1)
        b = gpu_ctx.sift(im, True)
    finally:
        gpu_ctx.set_option("sift_describe_gather", 0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.abs(a[2] - b[2]).max() < 2e-6


def _texture(rng, h, w):
    im = rng.random((h, w)).astype(np.float32)
    for _ in range(2):
        im = (im + np.roll(im, 1, 0) + np.roll(im, -1, 0) + np.roll(im, 1, 1) + np.roll(im, -1, 1)) / 5
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(max(4, h * w // 900)):
        cy, cx, s = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(1.5, 7)
        im += rng.uniform(-0.6, 0.6) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
    return ((im - im.min()) / (im.max() - im.min() + 1e-9) * 255).astype(np.uint8)


@pytest.mark.parametrize("shape,dbl", [((13, 13), False), ((13, 200), True), ((97, 131), True), ((64, 65), False), ((240, 127), True),
                                       ((33, 500), False), ((480, 640), False)])
def test_random_images_of_awkward_sizes_match_oracle(gpu_ctx, oracle_mod, shape, dbl):
    """Sizes that are not multiples of the 64-pixel blur tiles or of the 128x8 blocks, octaves that end at 13 pixels,
    windows that hang over every image border, saturated pixels."""
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    im = _texture(rng, *shape)
    im[: shape[0] // 7] = 255                                   # a saturated band: zero gradients inside, a hard edge below
    got = gpu_ctx.sift(im, dbl)
    want = oracle_mod.sift(im, dbl)
    assert len(got[0]) == len(want[0])
    if len(want[0]):
        compare_with_oracle(got, want, min_frac=0.99)
    n_oct = len(oracle_mod.sift_octave_dims(*shape, dbl))
    if n_oct:
        gauss, dog, _ = oracle_mod.sift_debug(im, dbl, n_oct - 1)
        assert np.array_equal(gpu_ctx.sift_plane(0, n_oct - 1, 0, 5), gauss[5])
        assert np.array_equal(gpu_ctx.sift_plane(0, n_oct - 1, 1, 4), dog[4])


def test_mixed_batch_sizes_reuse_the_context(gpu_ctx, gold):
    """Re-planning: different image sizes and batch sizes through one context, results independent of what ran before."""
    a = gold["bag0_crop_double/image"]
    b = gold["ex2_odd_double/image"]
    ra1 = gpu_ctx.sift(a, True)
    rb = gpu_ctx.sift(np.stack([b, b]), True)
    ra2 = gpu_ctx.sift(a, False)
    ra3 = gpu_ctx.sift(a, True)
    for x, y in zip(ra1, ra3):
        assert np.array_equal(x, y)
    assert np.array_equal(rb[0][2], rb[1][2]) and len(ra2[0]) < len(ra1[0])
