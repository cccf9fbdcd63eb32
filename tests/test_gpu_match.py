"""MATCH through the C ABI on a B200 vs the oracle (exact 2-NN in the reference's arithmetic): bit-exact rows,
distances and accepted set, in both the tensor-core mode and the exhaustive exact mode."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(ctx, capi, oracle, dbn, qn, ratio=0.8, modes=None):
    oidx, odist = oracle.match_2nn(dbn, qn)
    acc_o = (odist[:, 0] / odist[:, 1] < np.float32(ratio)) & (oidx[:, 1] >= 0)
    for mode in modes or (capi.MATCH_TENSOR, capi.MATCH_EXACT):
        r, d, a, st = ctx.match(qn, ratio, mode)
        assert np.array_equal(r, oidx), f"mode {mode}: rows differ at {np.nonzero((r != oidx).any(1))[0][:5]}"
        assert np.array_equal(d, odist), f"mode {mode}: distances differ"
        assert np.array_equal(a, acc_o)
    return st


def test_golden_fixture(gpu_ctx, golden):
    from moped_b200 import capi
    ctx = gpu_ctx
    ctx.db_upload(golden["db_desc"], golden["db_xyz"], golden["model_of_row"], len(golden["n_pts"]))
    for mode in (capi.MATCH_TENSOR, capi.MATCH_EXACT):
        r, d, a, st = ctx.match(golden["q_desc"], 0.8, mode)
        assert np.array_equal(r, golden["ann_idx"])            # the reference's kd-tree in exact mode (eps = 0)
        assert np.array_equal(d, golden["ann_dist"])
    # informational: agreement with the shipped approximate default (eps = 5) on the accepted matches
    acc = a.astype(bool)
    agree = (golden["ann5_idx"][acc, 0] == r[acc, 0]).mean()
    assert agree > 0.9


def test_small_case_both_modes(gpu_ctx, small_case, oracle_mod):
    from moped_b200 import capi
    c = small_case
    gpu_ctx.db_upload(c["dbn"], c["db"]["xyz"], c["db"]["model_of_row"], c["n_obj"])
    st = _check(gpu_ctx, capi, oracle_mod, c["dbn"], c["qn"])
    assert st[1] == c["qn"].shape[0]      # exact mode ran last: every query scanned exhaustively


@pytest.mark.parametrize("n_rows,q", [(1, 5), (2, 3), (3, 300), (127, 1), (129, 257), (1000, 255), (4097, 513)])
def test_ragged_sizes(gpu_ctx, oracle_mod, n_rows, q):
    """database not a multiple of the 128-row tile, queries not a multiple of 256, fewer rows than k, a single row."""
    from moped_b200 import capi, synth
    rng = np.random.default_rng(n_rows * 7 + q)
    dbn = oracle_mod.norm_rows(synth.sift_like(rng, n_rows))
    qn = oracle_mod.norm_rows(synth.sift_like(rng, q))
    gpu_ctx.db_upload(dbn, np.zeros((n_rows, 3), np.float32), np.zeros(n_rows, np.int32), 1)
    _check(gpu_ctx, capi, oracle_mod, dbn, qn)


def test_exact_ties_take_the_lower_row(gpu_ctx, oracle_mod):
    """duplicate descriptors: equal distances, the earlier row wins like a brute-force scan in row order"""
    from moped_b200 import capi, synth
    rng = np.random.default_rng(5)
    base = oracle_mod.norm_rows(synth.sift_like(rng, 300))
    dbn = np.concatenate([base, base[:50]])           # rows 300..349 duplicate rows 0..49
    qn = np.concatenate([base[:20], oracle_mod.norm_rows(synth.sift_like(rng, 44))])
    gpu_ctx.db_upload(dbn, np.zeros((len(dbn), 3), np.float32), np.zeros(len(dbn), np.int32), 1)
    _check(gpu_ctx, capi, oracle_mod, dbn, qn)
    r, d, a, _ = gpu_ctx.match(qn, 0.8, capi.MATCH_TENSOR)
    assert np.array_equal(r[:20, 0], np.arange(20)) and np.array_equal(r[:20, 1], np.arange(20) + 300) and (d[:20] == 0).all()
    assert not a[:20].any()          # 0/0 is not < ratio


def test_signed_descriptors(gpu_ctx, oracle_mod):
    """SURF-like signed descriptors: scores of either sign in the coarse pass"""
    from moped_b200 import capi
    rng = np.random.default_rng(9)
    dbn = oracle_mod.norm_rows(rng.normal(size=(3000, 128)).astype(np.float32))
    qn = oracle_mod.norm_rows((dbn[rng.integers(0, 3000, 200)] + rng.normal(0, 0.02, size=(200, 128))).astype(np.float32))
    gpu_ctx.db_upload(dbn, np.zeros((3000, 3), np.float32), np.zeros(3000, np.int32), 1)
    _check(gpu_ctx, capi, oracle_mod, dbn, qn)


def test_row_base_offsets_global_ids(gpu_ctx, oracle_mod):
    from moped_b200 import capi, synth
    rng = np.random.default_rng(2)
    dbn = oracle_mod.norm_rows(synth.sift_like(rng, 700)); qn = oracle_mod.norm_rows(synth.sift_like(rng, 90))
    gpu_ctx.db_upload(dbn, np.zeros((700, 3), np.float32), np.zeros(700, np.int32), 1, row_base=12345)
    r, d, a, _ = gpu_ctx.match(qn, 0.8, capi.MATCH_TENSOR)
    oidx, odist = oracle_mod.match_2nn(dbn, qn)
    assert np.array_equal(r, oidx + 12345) and np.array_equal(d, odist)


def test_full_size_properties(gpu_ctx):
    """BASELINE configs[1]: 100 objects / 100k descriptors, 2000 features. Too big for the CPU oracle in seconds:
    the tensor-core path must equal the GPU's own exhaustive exact scan bit for bit, distances must be
    reproducible from the rows in the reference's summation order, and planted features must find their row."""
    from moped_b200 import capi, synth
    db = synth.make_db(100, 1000)
    fr = synth.make_frame(db, 2000, n_visible=8)
    n = np.sqrt((db["desc"] ** 2).sum(1, dtype=np.float32)); dbn = (db["desc"] / n[:, None]).astype(np.float32)
    n = np.sqrt((fr["desc"] ** 2).sum(1, dtype=np.float32)); qn = (fr["desc"] / n[:, None]).astype(np.float32)
    gpu_ctx.db_upload(dbn, db["xyz"], db["model_of_row"], 100)
    rt, dt, at, st = gpu_ctx.match(qn, 0.8, capi.MATCH_TENSOR)
    re_, de, ae, _ = gpu_ctx.match(qn, 0.8, capi.MATCH_EXACT)
    assert np.array_equal(rt, re_) and np.array_equal(dt, de) and np.array_equal(at, ae)
    assert st[0] + st[1] == 2000
    assert (dt[:, 0] <= dt[:, 1]).all()
    for qi in range(0, 2000, 97):          # recompute in the reference's order: d = d + (q-p)*(q-p), fp32
        for k in range(2):
            p = dbn[rt[qi, k]]
            acc = np.float32(0)
            for t in (qn[qi] - p).astype(np.float32):
                acc = np.float32(acc + np.float32(t * t))
            assert acc == dt[qi, k]
    planted = fr["src_row"] >= 0
    assert (rt[planted, 0] == fr["src_row"][planted]).mean() > 0.99
    assert at[planted].mean() > 0.95 and at[~planted].mean() < 0.05


def test_errors_are_loud(gpu_ctx):
    from moped_b200 import capi
    ctx = capi.Context(0)
    with pytest.raises(capi.MopedCudaError):
        ctx.match(np.zeros((4, 128), np.float32))           # no database yet
    with pytest.raises(capi.MopedCudaError):
        ctx.db_upload(np.zeros((4, 64), np.float32), np.zeros((4, 3), np.float32), np.zeros(4, np.int32), 1)   # unsupported length
    ctx.close()
