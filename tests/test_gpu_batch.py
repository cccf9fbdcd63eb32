"""Frame batches through the C ABI (mc_process_frames*, BASELINE.json configs[4]): every frame of a batch must give
exactly the objects mc_process_frame gives for it alone — the batch only changes scheduling (one MATCH pass for all
queries, concurrent lanes after it), never results. Covers ragged and empty frames and every tuning setting."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def batch_case(oracle_mod):
    from moped_b200 import synth
    db = synth.make_db(30, 1000)
    sizes = [2000, 700, 0, 1500, 2000, 256, 1, 2000, 1234]
    frames = [synth.make_frame(db, q, n_visible=(4 if q >= 700 else 1), frame_id=10 + i) if q > 0 else None for i, q in enumerate(sizes)]
    dbn = oracle_mod.norm_rows(db["desc"])
    qn = [oracle_mod.norm_rows(f["desc"]) if f is not None else np.zeros((0, 128), np.float32) for f in frames]
    xy = [f["xy"] if f is not None else np.zeros((0, 2), np.float32) for f in frames]
    img = [f["image_idx"] if f is not None else np.zeros(0, np.int32) for f in frames]
    return dict(db=db, dbn=dbn, frames=frames, qn=qn, xy=xy, img=img, sizes=sizes)


@pytest.fixture()
def ctx30(gpu_ctx, batch_case):
    from moped_b200 import synth
    c = batch_case
    gpu_ctx.db_upload(c["dbn"], c["db"]["xyz"], c["db"]["model_of_row"], 30)
    gpu_ctx.set_cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    gpu_ctx.set_tuning(8, 8, 1)
    return gpu_ctx


def _single(ctx, c):
    out = []
    for q, xy, img in zip(c["qn"], c["xy"], c["img"]):
        if len(q) == 0:
            out.append(dict(model=np.zeros(0, np.int32), pose=np.zeros((0, 7), np.float32), score=np.zeros(0, np.float32)))
        else:
            out.append(ctx.process_frame(q, xy, img, max_objects=64))
    return out


def _same(a, b):
    assert len(a) == len(b)
    for f, (x, y) in enumerate(zip(a, b)):
        assert np.array_equal(x["model"], y["model"]), f
        assert np.array_equal(x["pose"], y["pose"]), f            # bit-identical: same kernels, same seeds
        assert np.array_equal(x["score"], y["score"]), f


@pytest.mark.parametrize("lanes,warps,chunks", [(1, 8, 1), (4, 2, 2), (8, 1, 3), (16, 4, 16)])
def test_batch_equals_single_frames(ctx30, batch_case, lanes, warps, chunks):
    c = batch_case
    ctx30.set_tuning(8, 8, 1)
    single = _single(ctx30, c)
    assert sum(len(s["model"]) for s in single) >= 12          # the planted objects are found
    ctx30.set_tuning(lanes, warps, chunks)
    fo = np.concatenate([[0], np.cumsum(c["sizes"])]).astype(np.int32)
    batch = ctx30.process_frames(np.concatenate(c["qn"]), np.concatenate(c["xy"]), np.concatenate(c["img"]), fo, max_objects=64)
    _same(batch, single)
    for f, b in enumerate(batch):
        assert b["info"][0] == len(b["model"]) and b["info"][1] == 0
        if c["sizes"][f] == 0:
            assert b["info"].tolist() == [0, 0, 0, 0]
    ctx30.set_tuning(8, 8, 1)


def test_batch_objects_match_ground_truth(ctx30, batch_case):
    c = batch_case
    fo = np.concatenate([[0], np.cumsum(c["sizes"])]).astype(np.int32)
    batch = ctx30.process_frames(np.concatenate(c["qn"]), np.concatenate(c["xy"]), np.concatenate(c["img"]), fo, max_objects=64)
    for f, fr in enumerate(c["frames"]):
        if fr is None or c["sizes"][f] < 700:
            continue
        assert sorted(batch[f]["model"].tolist()) == sorted(fr["gt_model"].tolist()), f
        for m, p in zip(batch[f]["model"], batch[f]["pose"]):
            g = fr["gt_pose"][list(fr["gt_model"]).index(m)]
            assert np.abs(p[4:] - g[4:]).max() < 0.01


def test_batch_device_entry_and_frame_range(ctx30, batch_case):
    """mc_process_frames_dev (queries resident) and mc_process_frames_matched_dev on a frame sub-range (the
    multi-GPU split of the stages after MATCH) give the same frames."""
    import torch
    c = batch_case
    dev = torch.device("cuda", 0)
    fo = np.concatenate([[0], np.cumsum(c["sizes"])]).astype(np.int32)
    Q = int(fo[-1])
    q = torch.from_numpy(np.concatenate(c["qn"])).to(dev)
    xy = torch.from_numpy(np.concatenate(c["xy"])).to(dev)
    img = torch.from_numpy(np.concatenate(c["img"])).to(dev)
    ref = ctx30.process_frames_dev(q.data_ptr(), xy.data_ptr(), img.data_ptr(), fo, max_objects=64)
    _same(ref, _single(ctx30, c))
    # matched entry: nearest neighbours of all queries from mc_match_dev, then frames [3, 8)
    p = ctx30.default_params()
    nn_row = torch.empty((Q, 2), dtype=torch.int32, device=dev)
    nn_dist = torch.empty((Q, 2), dtype=torch.float32, device=dev)
    acc = torch.empty((Q,), dtype=torch.uint8, device=dev)
    ctx30.match_dev(q.data_ptr(), Q, p.match_ratio, p.match_mode, nn_row.data_ptr(), nn_dist.data_ptr(), acc.data_ptr())
    nf, mo = 5, 64
    info = torch.zeros((nf, 4), dtype=torch.int32, device=dev)
    om = torch.zeros((nf, mo), dtype=torch.int32, device=dev)
    op = torch.zeros((nf, mo, 7), dtype=torch.float32, device=dev)
    osc = torch.zeros((nf, mo), dtype=torch.float32, device=dev)
    ctx30.process_frames_matched_dev(nn_row.data_ptr(), acc.data_ptr(), xy.data_ptr(), img.data_ptr(), fo, 3, 8, p, mo,
                                     info.data_ptr(), om.data_ptr(), op.data_ptr(), osc.data_ptr())
    ctx30.synchronize()
    torch.cuda.synchronize()
    info, om, op, osc = info.cpu().numpy(), om.cpu().numpy(), op.cpu().numpy(), osc.cpu().numpy()
    for s in range(nf):
        k = int(info[s, 0])
        assert np.array_equal(om[s, :k], ref[3 + s]["model"])
        assert np.array_equal(op[s, :k], ref[3 + s]["pose"])
        assert np.array_equal(osc[s, :k], ref[3 + s]["score"])


def test_batch_argument_errors(ctx30, batch_case):
    from moped_b200 import capi
    c = batch_case
    with pytest.raises(capi.MopedCudaError):
        ctx30.process_frames(c["qn"][0], c["xy"][0], c["img"][0], np.array([0, 100, 50], np.int32))
    with pytest.raises(capi.MopedCudaError):
        ctx30.process_frames(c["qn"][0], c["xy"][0], c["img"][0], np.array([5, 100], np.int32))
    with pytest.raises(capi.MopedCudaError):
        ctx30.set_tuning(65, 0, 0)


def test_sharded_batch_equals_single_context(ctx30, batch_case):
    """The N = 2 data path of bench.py on one GPU: two contexts hold the two object shards (+ the global coord3D /
    model tables), each matches ALL queries of the batch against its shard, the (row, distance) pairs are stacked
    (= the all-gather) and merged, then "rank" r runs CLUSTER..FILTER2 for its half of the frames and writes its
    result block; the unpacked blocks must equal the single-context batch bit for bit."""
    import torch
    from moped_b200 import capi, synth
    from moped_b200.sharding import ResultBlock, frame_range, shard_objects
    c = batch_case
    sizes = [2000, 700, 0, 1500, 2000, 256, 1, 2000]          # 8 frames -> 4 per rank
    fo = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    Q = int(fo[-1])
    qn, xy, img = np.concatenate(c["qn"][:8]), np.concatenate(c["xy"][:8]), np.concatenate(c["img"][:8])
    ref = ctx30.process_frames(qn, xy, img, fo, max_objects=32)
    dev = torch.device("cuda", 0)
    dq, dxy, dimg = torch.from_numpy(qn).to(dev), torch.from_numpy(xy).to(dev), torch.from_numpy(img).to(dev)
    world, MO = 2, 32
    rows_all = torch.empty((world, Q, 2), dtype=torch.int32, device=dev)
    dist_all = torch.empty((world, Q, 2), dtype=torch.float32, device=dev)
    acc = torch.empty(Q, dtype=torch.uint8, device=dev)
    ctxs = []
    for r, (o0, o1, r0, r1) in enumerate(shard_objects(c["db"]["n_pts"], world)):
        cx = capi.Context(0)
        cx.db_upload(c["dbn"][r0:r1], c["db"]["xyz"][r0:r1], c["db"]["model_of_row"][r0:r1], 30, row_base=r0)
        cx.db_set_global_tables(c["db"]["xyz"], c["db"]["model_of_row"], 30)
        cx.set_cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
        cx.set_tuning(4, 2, 1)
        cx.match_dev(dq.data_ptr(), Q, 0.8, capi.MATCH_TENSOR, rows_all[r].data_ptr(), dist_all[r].data_ptr(), acc.data_ptr())
        cx.synchronize()
        ctxs.append(cx)
    blocks = []
    p = ctxs[0].default_params()
    for r, cx in enumerate(ctxs):
        nn_row = torch.empty((Q, 2), dtype=torch.int32, device=dev); nn_dist = torch.empty((Q, 2), dtype=torch.float32, device=dev)
        a = torch.empty(Q, dtype=torch.uint8, device=dev)
        cx.match_merge_dev(rows_all.data_ptr(), dist_all.data_ptr(), world, Q, 0.8, nn_row.data_ptr(), nn_dist.data_ptr(), a.data_ptr())
        lo, hi = frame_range(8, world, r)
        blk = ResultBlock(hi - lo, MO)
        buf = torch.zeros(blk.words, dtype=torch.int32, device=dev)
        cx.process_frames_matched_dev(nn_row.data_ptr(), a.data_ptr(), dxy.data_ptr(), dimg.data_ptr(), fo, lo, hi, p, MO,
                                      buf.data_ptr() + 4 * blk.o_info, buf.data_ptr() + 4 * blk.o_model, buf.data_ptr() + 4 * blk.o_pose,
                                      buf.data_ptr() + 4 * blk.o_score)
        cx.synchronize()
        blocks.append(buf.cpu().numpy())
    got = blk.unpack(np.stack(blocks))
    _same(got, ref)
    for cx in ctxs:
        cx.close()


@pytest.mark.parametrize("world,exact", [(2, 0), (3, 0), (2, 1)])
def test_cluster_partitioned_frame_equals_single_context(ctx30, batch_case, world, exact):
    """north_star: "RANSAC work is distributed by cluster". mc_process_frame_sharded_dev on `world` contexts of one GPU, each standing for
    a rank: all run compaction / CLUSTER / FILTER on the same nearest neighbours and the POSE / POSE2 tasks of their own clusters
    (cluster % world == rank); the records are copied between the ranks' exchange buffers like the in-place all-gather does. Objects
    (model, pose, score) and the frame info equal the single-context frame bit for bit on every rank — a task's random stream depends
    on its index alone — also in exact-order mode."""
    import torch
    from moped_b200 import capi, synth
    c = batch_case
    dev = torch.device("cuda", 0)
    MO = 32
    ctxs = []
    for r in range(world):
        cx = capi.Context(0)
        cx.db_upload(c["dbn"], c["db"]["xyz"], c["db"]["model_of_row"], 30)
        cx.set_cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
        cx.set_tuning(4, 2 + r, 1)                                   # the first-round width must not matter either
        cx.set_option("pose_exact_order", exact)
        ctxs.append(cx)
    ctx30.set_option("pose_exact_order", exact)
    try:
        n_checked = 0
        for fi in (0, 3, 5, 8):
            qn, xy, img = c["qn"][fi], c["xy"][fi], c["img"][fi]
            Q = len(qn)
            ref = ctx30.process_frame(qn, xy, img, max_objects=MO)
            dq, dxy, dimg = torch.from_numpy(qn).to(dev), torch.from_numpy(xy).to(dev), torch.from_numpy(img).to(dev)
            nn_row = torch.empty((Q, 2), dtype=torch.int32, device=dev); nn_dist = torch.empty((Q, 2), dtype=torch.float32, device=dev)
            acc = torch.empty(Q, dtype=torch.uint8, device=dev)
            p = ctxs[0].default_params()
            ctxs[0].match_dev(dq.data_ptr(), Q, p.match_ratio, p.match_mode, nn_row.data_ptr(), nn_dist.data_ptr(), acc.data_ptr())
            ctxs[0].synchronize()
            slot = ctxs[0].frame_shard_slot_bytes(Q, p)
            assert slot > 0 and slot % 256 == 0
            exch = [torch.zeros((world, slot), dtype=torch.uint8, device=dev) for _ in ctxs]      # every rank's own buffer, as over NCCL
            outs = [dict(info=torch.zeros(4, dtype=torch.int32, device=dev), model=torch.zeros(MO, dtype=torch.int32, device=dev),
                         pose=torch.zeros((MO, 7), dtype=torch.float32, device=dev), score=torch.zeros(MO, dtype=torch.float32, device=dev)) for _ in ctxs]
            for phase in range(3):
                for r, cx in enumerate(ctxs):
                    o = outs[r]
                    cx.process_frame_sharded_dev(phase, nn_row.data_ptr(), acc.data_ptr(), dxy.data_ptr(), dimg.data_ptr(), Q, p, r, world,
                                                 exch[r].data_ptr(), MO, o["info"].data_ptr(), o["model"].data_ptr(), o["pose"].data_ptr(), o["score"].data_ptr())
                for cx in ctxs:
                    cx.synchronize()
                for r in range(world):                                # the in-place all-gather: slot s of every buffer <- rank s's record
                    for s_ in range(world):
                        if s_ != r:
                            exch[r][s_].copy_(exch[s_][s_])
                torch.cuda.synchronize()
            for r, o in enumerate(outs):
                info = o["info"].cpu().numpy()
                n = int(info[0])
                assert info[1] == 0 and n == len(ref["model"]), (fi, r, info, len(ref["model"]))
                assert np.array_equal(o["model"].cpu().numpy()[:n], ref["model"]), (fi, r)
                assert np.array_equal(o["pose"].cpu().numpy()[:n], ref["pose"]), (fi, r)
                assert np.array_equal(o["score"].cpu().numpy()[:n], ref["score"]), (fi, r)
            n_checked += len(ref["model"])
        assert n_checked >= 6
    finally:
        ctx30.set_option("pose_exact_order", 0)
        for cx in ctxs:
            cx.close()


def test_batch_graph_equals_per_lane_graphs(ctx30, batch_case):
    """mc_set_option("batch_graph"): the stage chains of all frames of a call captured into ONE CUDA graph (a branch per lane) give what
    the per-lane path gives — on the capture call, on replays, after a change of the batch (another key), with fewer lanes than frames
    (several frames per branch), with empty frames, and through the device-resident entry."""
    import torch
    c = batch_case
    fo = np.concatenate([[0], np.cumsum(c["sizes"])]).astype(np.int32)
    q, xy, img = np.concatenate(c["qn"]), np.concatenate(c["xy"]), np.concatenate(c["img"])
    order2 = [8, 0, 3, 1]
    q2 = np.concatenate([c["qn"][i] for i in order2]); xy2 = np.concatenate([c["xy"][i] for i in order2]); img2 = np.concatenate([c["img"][i] for i in order2])
    fo2 = np.concatenate([[0], np.cumsum([c["sizes"][i] for i in order2])]).astype(np.int32)
    try:
        ctx30.set_option("batch_graph", 0)
        ctx30.set_tuning(16, 4, 1)
        want = ctx30.process_frames(q, xy, img, fo, max_objects=64)
        want2 = ctx30.process_frames(q2, xy2, img2, fo2, max_objects=64)
        assert sum(len(w["model"]) for w in want) >= 12
        ctx30.set_option("batch_graph", 1)
        for lanes in (16, 3):
            ctx30.set_tuning(lanes, 4, 1)
            l0 = ctx30.launches
            for rep in range(4):                                   # allocation call, capture call, replays
                _same(ctx30.process_frames(q, xy, img, fo, max_objects=64), want)
            assert ctx30.launches > l0
            _same(ctx30.process_frames(q2, xy2, img2, fo2, max_objects=64), want2)
            _same(ctx30.process_frames(q2, xy2, img2, fo2, max_objects=64), want2)
            _same(ctx30.process_frames(q, xy, img, fo, max_objects=64), want)
        dev = torch.device("cuda", 0)
        dq, dxy, dimg = torch.from_numpy(q).to(dev), torch.from_numpy(xy).to(dev), torch.from_numpy(img).to(dev)
        for rep in range(3):
            _same(ctx30.process_frames_dev(dq.data_ptr(), dxy.data_ptr(), dimg.data_ptr(), fo, ctx30.default_params(), 64), want)
    finally:
        ctx30.set_option("batch_graph", 1)
        ctx30.set_tuning(8, 8, 1)


def test_frame_graphs_equal_eager_launches(ctx30, batch_case):
    """mc_process_frames replays one CUDA graph per frame for the stages after MATCH (captured per lane and per
    feature-count bucket). Replays, re-captures after a bucket change and the eager path give identical results,
    also when the same context then processes a second, different batch."""
    c = batch_case
    order = [0, 4, 7, 3, 8, 1, 5, 6, 2, 0, 4, 7, 7, 4]                  # repeated buckets on few lanes -> replays
    q = np.concatenate([c["qn"][i] for i in order])
    xy = np.concatenate([c["xy"][i] for i in order])
    img = np.concatenate([c["img"][i] for i in order])
    fo = np.concatenate([[0], np.cumsum([c["sizes"][i] for i in order])]).astype(np.int32)
    try:
        ctx30.set_tuning(2, 4, 1)
        ctx30.set_option("frame_graphs", 0)
        eager = ctx30.process_frames(q, xy, img, fo, max_objects=64)
        l0 = ctx30.launches
        eager2 = ctx30.process_frames(q, xy, img, fo, max_objects=64)
        n_eager = ctx30.launches - l0
        ctx30.set_option("frame_graphs", 1)
        for _ in range(3):                                               # first pass captures, later passes only replay
            l0 = ctx30.launches
            graph = ctx30.process_frames(q, xy, img, fo, max_objects=64)
            n_graph = ctx30.launches - l0
            _same(graph, eager)
        _same(eager2, eager)
        assert n_graph == n_eager                                        # replays are counted kernel by kernel
        assert sum(len(g["model"]) for g in graph) >= 20
        # a different batch on the same context (other buckets first)
        rev = order[::-1]
        q2 = np.concatenate([c["qn"][i] for i in rev]); xy2 = np.concatenate([c["xy"][i] for i in rev]); img2 = np.concatenate([c["img"][i] for i in rev])
        fo2 = np.concatenate([[0], np.cumsum([c["sizes"][i] for i in rev])]).astype(np.int32)
        g2 = ctx30.process_frames(q2, xy2, img2, fo2, max_objects=64)
        _same(g2, eager[::-1])
    finally:
        ctx30.set_option("frame_graphs", 1)
        ctx30.set_tuning(8, 8, 1)
