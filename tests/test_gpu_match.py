"""MATCH through the C ABI on a B200 vs the oracle (exact 2-NN in the reference's arithmetic): bit-exact rows,
distances and accepted set, in both the tensor-core mode and the exhaustive exact mode."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(ctx, capi, oracle, dbn, qn, ratio=0.8, modes=None):
    """every mode — the cascade that starts with the 8-bit integer coarse pass (the default), the fp16 coarse pass alone, the
    exhaustive scan — returns the oracle's rows, distances and accepted set bit for bit"""
    oidx, odist = oracle.match_2nn(dbn, qn)
    acc_o = (odist[:, 0] / odist[:, 1] < np.float32(ratio)) & (oidx[:, 1] >= 0)
    for mode in modes or (capi.MATCH_TENSOR, capi.MATCH_EXACT):
        for kind in ((1, 0) if mode == capi.MATCH_TENSOR else (1,)):
            ctx.set_option("match_coarse_kind", kind)
            try:
                r, d, a, st = ctx.match(qn, ratio, mode)
                tiers = ctx.match_tier_stats()
            finally:
                ctx.set_option("match_coarse_kind", 1)
            assert np.array_equal(r, oidx), f"mode {mode} kind {kind}: rows differ at {np.nonzero((r != oidx).any(1))[0][:5]}"
            assert np.array_equal(d, odist), f"mode {mode} kind {kind}: distances differ"
            assert np.array_equal(a, acc_o)
            assert tiers[0] == len(qn) and tiers[1] + tiers[2] + tiers[3] == len(qn), tiers
            if mode == capi.MATCH_TENSOR and ctx.D == 128:
                assert st[1] == tiers[3] and (kind == 1 or tiers[1] == 0), (st, tiers)
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


def test_awkward_value_ranges_for_the_8_bit_pass(gpu_ctx, oracle_mod):
    """What the 8-bit quantisation has to get right or hand to the next tier: query tiles of mixed signedness (128 non-negative
    queries next to 128 signed ones: u8 and s8 A operands in one launch), an all-zero query, queries of very different norms
    (the scale is per query), a database row with one huge element (it sets the database-wide scale, everything else quantises
    coarsely and the certificate has to refuse), un-normalised rows."""
    from moped_b200 import capi, synth
    rng = np.random.default_rng(31)
    db = synth.sift_like(rng, 5000)
    db[17] *= 3.0
    qpos = synth.sift_like(rng, 128)
    qneg = oracle_mod.norm_rows(rng.normal(size=(128, 128)).astype(np.float32))
    qmix = np.concatenate([qpos, qneg, 40.0 * synth.sift_like(rng, 60), 1e-3 * synth.sift_like(rng, 60), np.zeros((1, 128), np.float32),
                           db[rng.integers(0, 5000, 100)] + rng.normal(0, 0.01, size=(100, 128)).astype(np.float32)]).astype(np.float32)
    gpu_ctx.db_upload(db, np.zeros((5000, 3), np.float32), np.zeros(5000, np.int32), 1)
    _check(gpu_ctx, capi, oracle_mod, db, qmix)
    spiky = db.copy()
    spiky[100, 5] = 50.0                               # one outlier element: database-wide scale 255 / 50
    gpu_ctx.db_upload(spiky, np.zeros((5000, 3), np.float32), np.zeros(5000, np.int32), 1)
    _check(gpu_ctx, capi, oracle_mod, spiky, qmix[:300])
    tiers = gpu_ctx.match_tier_stats()                 # (exact mode ran last)
    gpu_ctx.set_option("match_coarse_kind", 1)
    gpu_ctx.match(qmix[:300], 0.8, capi.MATCH_TENSOR)
    tiers = gpu_ctx.match_tier_stats()
    assert tiers[1] < 300                              # the 8-bit pass could not certify everything here; the fp16 pass picked it up
    signed_db = oracle_mod.norm_rows(rng.normal(size=(3000, 128)).astype(np.float32))
    gpu_ctx.db_upload(signed_db, np.zeros((3000, 3), np.float32), np.zeros(3000, np.int32), 1)
    _check(gpu_ctx, capi, oracle_mod, signed_db, qmix)


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
        ctx.db_upload(np.zeros((4, 5000), np.float32), np.zeros((4, 3), np.float32), np.zeros(4, np.int32), 1)   # beyond the longest supported length
    ctx.close()


@pytest.mark.parametrize("d", [64, 36, 1, 200])
def test_other_descriptor_lengths(oracle_mod, d):
    """MATCH_ANN_CPU(int DescriptorSize, ...) works for any length (MATCH_ANN_CPU.hpp:113; SURF is 64-d): lengths other than
    128 take the exhaustive exact scan whatever mode is asked for — rows, distances and accepted set equal the oracle's."""
    from moped_b200 import capi
    rng = np.random.default_rng(100 + d)
    dbn = oracle_mod.norm_rows(rng.normal(size=(2500, d)).astype(np.float32))
    qn = oracle_mod.norm_rows(np.concatenate([dbn[rng.integers(0, 2500, 150)] + rng.normal(0, 0.05, size=(150, d)),
                                              rng.normal(size=(61, d))]).astype(np.float32))
    ctx = capi.Context(0)
    try:
        ctx.db_upload(dbn, np.zeros((2500, 3), np.float32), np.zeros(2500, np.int32), 1)
        st = _check(ctx, capi, oracle_mod, dbn, qn)
        assert st[1] == len(qn)
        r, dd, a, st = ctx.match(qn, 0.8, capi.MATCH_TENSOR)
        assert st[0] == 0 and st[1] == len(qn)              # reported as what it is: an exhaustive scan
    finally:
        ctx.close()


def test_second_context_in_one_process(oracle_mod, small_case):
    """Kernel attributes (opt-in shared memory of k_match_coarse / k_meanshift) are set per context at mc_create, not behind a
    process-wide flag: a context created after another one has already run works (with several GPUs: on ANY device)."""
    from moped_b200 import capi
    c = small_case
    first = capi.Context(0)
    first.db_upload(c["dbn"], c["db"]["xyz"], c["db"]["model_of_row"], c["n_obj"])
    first.match(c["qn"][:300], 0.8, capi.MATCH_TENSOR)
    import torch
    dev = torch.cuda.device_count() - 1                      # the last device: a different one on a multi-GPU box
    second = capi.Context(dev)
    try:
        second.db_upload(c["dbn"], c["db"]["xyz"], c["db"]["model_of_row"], c["n_obj"])
        _check(second, capi, oracle_mod, c["dbn"], c["qn"][:300], modes=(capi.MATCH_TENSOR,))
        from moped_b200 import synth
        second.set_cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
        out = second.process_frame(c["qn"], c["fr"]["xy"], c["fr"]["image_idx"])
        assert sorted(out["model"].tolist()) == sorted(c["fr"]["gt_model"].tolist())
    finally:
        second.close(); first.close()


def test_metric_configuration_1m_rows(oracle_mod):
    """The configuration bench.py times (BASELINE metric): 1000 objects / 1 M descriptors, a batch of 64 x 2000 = 128 000 queries in
    ONE mc_match_dev pass (10 DB splits x k=4 = 40 coarse candidates per query, another certificate regime than a single frame's 18
    splits). 256 sampled queries' rows AND distances equal the oracle's bit for bit (256 x 1 M is seconds on the host); every query is
    accounted for (certified + fallback == Q); the same batch against the database sharded over two contexts, merged by
    mc_match_merge_dev, gives the identical result for ALL 128 000 queries."""
    import torch
    from moped_b200 import capi, synth
    from moped_b200.sharding import shard_objects
    B, Q = 64, 2000
    db = synth.make_db(1000, 1000)
    dbn = oracle_mod.norm_rows(db["desc"])
    qn = np.concatenate([oracle_mod.norm_rows(synth.make_frame(db, Q, n_visible=8, frame_id=i)["desc"]) for i in range(B)])
    QT = len(qn)
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(qn).to(dev)

    def run(ctx):
        row = torch.full((QT, 2), -1, dtype=torch.int32, device=dev)
        dist = torch.zeros((QT, 2), dtype=torch.float32, device=dev)
        acc = torch.zeros(QT, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ctx.match_dev(d_q.data_ptr(), QT, 0.8, capi.MATCH_TENSOR, row.data_ptr(), dist.data_ptr(), acc.data_ptr())
        st = ctx.match_last_stats()
        tiers = ctx.match_tier_stats()
        assert tiers[0] == QT and tiers[1] + tiers[2] + tiers[3] == QT and tiers[3] == st[1], (tiers, st)
        ctx.synchronize()
        return row, dist, acc, st

    ctx = capi.Context(0)
    try:
        ctx.db_upload(dbn, db["xyz"], db["model_of_row"], 1000)
        row, dist, acc, st = run(ctx)
        tiers = ctx.match_tier_stats()
        assert tiers[1] >= 0.98 * QT, f"the 8-bit pass certified only {tiers[1]} of {QT} queries"
        ctx.set_option("match_coarse_kind", 0)         # the fp16 pass alone: the same bits for all 128 000 queries
        row0, dist0, acc0, st0 = run(ctx)
        assert torch.equal(row0, row) and torch.equal(dist0, dist) and torch.equal(acc0, acc)
    finally:
        ctx.close()
    assert st[0] + st[1] == QT and st[2] >= 32, st
    assert st[0] >= 0.99 * QT, f"only {st[0]} of {QT} queries certified by the coarse pass"
    r, d, a = row.cpu().numpy(), dist.cpu().numpy(), acc.cpu().numpy().astype(bool)
    rng = np.random.default_rng(7)
    pick = np.sort(rng.choice(QT, 256, replace=False))
    oidx, odist = oracle_mod.match_2nn(dbn, qn[pick])
    assert np.array_equal(r[pick], oidx), np.nonzero((r[pick] != oidx).any(1))[0][:5]
    assert np.array_equal(d[pick], odist)
    assert np.array_equal(a[pick], odist[:, 0] / odist[:, 1] < np.float32(0.8))
    # two object shards on two contexts, merged
    shards = shard_objects(db["n_pts"], 2)
    rows_all = torch.empty((2, QT, 2), dtype=torch.int32, device=dev)
    dist_all = torch.empty((2, QT, 2), dtype=torch.float32, device=dev)
    ctxs = [capi.Context(0) for _ in shards]
    try:
        tot = np.zeros(2, np.int64)
        for k, (cx, (o0, o1, r0, r1)) in enumerate(zip(ctxs, shards)):
            cx.db_upload(dbn[r0:r1], db["xyz"][r0:r1], db["model_of_row"][r0:r1], 1000, row_base=r0)
            rk, dk, _, stk = run(cx)
            assert stk[0] + stk[1] == QT
            tot += stk[:2]
            rows_all[k].copy_(rk); dist_all[k].copy_(dk)
        m_row = torch.empty((QT, 2), dtype=torch.int32, device=dev)
        m_dist = torch.empty((QT, 2), dtype=torch.float32, device=dev)
        m_acc = torch.empty(QT, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ctxs[0].match_merge_dev(rows_all.data_ptr(), dist_all.data_ptr(), 2, QT, 0.8, m_row.data_ptr(), m_dist.data_ptr(), m_acc.data_ptr())
        ctxs[0].synchronize()
    finally:
        for cx in ctxs:
            cx.close()
    assert np.array_equal(m_row.cpu().numpy(), r) and np.array_equal(m_dist.cpu().numpy(), d)
    assert np.array_equal(m_acc.cpu().numpy().astype(bool), a)
