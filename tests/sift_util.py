"""Keypoint-list comparison shared by the SIFT oracle and CUDA tests."""
import numpy as np
from scipy.spatial import cKDTree


def match_keypoints(xy_a, desc_a, xy_b, desc_b, tol_px=0.02):
    """For every keypoint of A the keypoint of B at the same place (within tol_px) with the closest descriptor.
    Returns (index into B or -1, max-abs descriptor difference). Several keypoints may share one location (one per
    orientation peak), hence the descriptor tie-break."""
    idx = np.full(len(xy_a), -1, np.int64)
    dd = np.full(len(xy_a), np.inf, np.float32)
    if len(xy_a) == 0 or len(xy_b) == 0:
        return idx, dd
    tree = cKDTree(xy_b)
    for i, nb in enumerate(tree.query_ball_point(xy_a, tol_px)):
        for j in nb:
            d = np.abs(desc_a[i] - desc_b[j]).max()
            if d < dd[i]:
                dd[i], idx[i] = d, j
    return idx, dd


def assert_same_keypoints(xy, desc, gxy, gdesc, min_matched=0.99, tol_px=0.02, tol_desc=5e-3, frac_tight=0.98, tight=1e-3):
    """Gate used against the -ffast-math reference: extrema sitting exactly on a threshold may flip, so at least
    `min_matched` of BOTH lists must pair up 1:1; paired descriptors within tol_desc, `frac_tight` of them within `tight`."""
    n, m = len(xy), len(gxy)
    assert abs(n - m) <= max(2, 0.01 * m), (n, m)
    ia, da = match_keypoints(xy, desc, gxy, gdesc, tol_px)
    ib, _ = match_keypoints(gxy, gdesc, xy, desc, tol_px)
    ok = ia >= 0
    assert ok.mean() >= min_matched and (ib >= 0).mean() >= min_matched, (ok.mean(), (ib >= 0).mean())
    assert len(np.unique(ia[ok])) == ok.sum()                       # 1:1
    good = da[ok]
    assert (good < tol_desc).mean() >= min_matched, np.sort(good)[-5:]
    assert (good < tight).mean() >= frac_tight, (good < tight).mean()
    if n == m and ok.all():
        return bool(np.array_equal(ia, np.arange(n)))              # same order too
    return False
