"""N > 1 host logic on CPU: object sharding + the all-gather exchange + the merge rule, world_size 2 over gloo.
Each rank matches all queries against ITS shard (with the oracle — there is no GPU here), the per-query
(global row, distance) pairs are all-gathered and merged exactly as k_match_merge does (smaller distance, then
smaller global row id); the result must equal the single-shard answer bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def merge_top2(rows_all, dist_all):
    """rows_all/dist_all: [shards, Q, 2] -> [Q, 2] each (test-side restatement of the merge rule)."""
    S, Q, _ = rows_all.shape
    rows = rows_all.transpose(1, 0, 2).reshape(Q, 2 * S)
    dist = dist_all.transpose(1, 0, 2).reshape(Q, 2 * S)
    out_r = np.empty((Q, 2), np.int32); out_d = np.empty((Q, 2), np.float32)
    for q in range(Q):
        valid = rows[q] >= 0
        order = sorted(np.nonzero(valid)[0], key=lambda j: (dist[q, j], rows[q, j]))
        for k in range(2):
            out_r[q, k] = rows[q, order[k]] if k < len(order) else -1
            out_d[q, k] = dist[q, order[k]] if k < len(order) else np.inf
    return out_r, out_d


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from moped_b200 import synth
    from moped_b200.sharding import shard_objects
    from oracle import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    db = synth.make_db(7, 150, seed=3, ragged=True)
    fr = synth.make_frame(db, 200, n_visible=3, pts_visible=30, seed=3)
    dbn, qn = oracle.norm_rows(db["desc"]), oracle.norm_rows(fr["desc"])
    o0, o1, r0, r1 = shard_objects(db["n_pts"], world)[rank]
    idx, d = oracle.match_2nn(dbn[r0:r1], qn)
    idx = np.where(idx >= 0, idx + r0, -1).astype(np.int32)
    rows_all = [torch.empty((len(qn), 2), dtype=torch.int32) for _ in range(world)]
    dist_all = [torch.empty((len(qn), 2), dtype=torch.float32) for _ in range(world)]
    dist.all_gather(rows_all, torch.from_numpy(idx))
    dist.all_gather(dist_all, torch.from_numpy(d))
    mr, md = merge_top2(torch.stack(rows_all).numpy(), torch.stack(dist_all).numpy())
    fidx, fd = oracle.match_2nn(dbn, qn)
    q.put((rank, bool(np.array_equal(mr, fidx)), bool(np.array_equal(md, fd)), (r0, r1)))
    dist.destroy_process_group()


def test_two_rank_sharded_match_equals_single_shard():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] and r[2] for r in res), res
    spans = sorted(r[3] for r in res)
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0]


def test_shard_objects_partitions_everything():
    from moped_b200.sharding import shard_objects
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        n_pts = rng.integers(600, 3000, size=1000)
        sh = shard_objects(n_pts, world)
        assert sh[0][0] == 0 and sh[-1][1] == 1000 and sh[-1][3] == n_pts.sum()
        for a, b in zip(sh[:-1], sh[1:]):
            assert a[1] == b[0] and a[3] == b[2]
        sizes = np.array([s[3] - s[2] for s in sh])
        assert sizes.max() - sizes.min() <= 3000 * 2


def test_synth_is_deterministic():
    from moped_b200 import synth
    a, b = synth.make_db(3, 50, seed=9), synth.make_db(3, 50, seed=9)
    assert np.array_equal(a["desc"], b["desc"]) and np.array_equal(a["xyz"], b["xyz"])
    assert np.allclose(np.linalg.norm(a["desc"], axis=1), 1.0, atol=1e-5)
    fa, fb = synth.make_frame(a, 100, n_visible=2, pts_visible=20), synth.make_frame(b, 100, n_visible=2, pts_visible=20)
    assert np.array_equal(fa["desc"], fb["desc"]) and len(fa["desc"]) == 100


def _worker_frames(rank, world, port, q):
    """Frame partition of a batch + the all-gather of the per-rank result blocks (what bench.py does over NCCL)."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from moped_b200.sharding import ResultBlock, frame_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, MO = 6, 5
    lo, hi = frame_range(B, world, rank)
    blk = ResultBlock(hi - lo, MO)
    buf = np.zeros(blk.words, np.int32)
    # frame f "finds" f % 4 objects: model id 100 + f, pose = f + slot/10, score = f
    for s, f in enumerate(range(lo, hi)):
        k = f % 4
        buf[blk.o_info + 4 * s:blk.o_info + 4 * s + 4] = [k, 0, 10 * f, f]
        buf[blk.o_model + MO * s:blk.o_model + MO * s + k] = 100 + f
        buf[blk.o_score + MO * s:blk.o_score + MO * s + k] = np.full(k, f, np.float32).view(np.int32)
        buf[blk.o_pose + 7 * MO * s:blk.o_pose + 7 * MO * s + 7 * k] = (f + np.arange(7 * k) // 7 / 10).astype(np.float32).view(np.int32)
    out = [torch.empty(blk.words, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(out, torch.from_numpy(buf))
    frames = blk.unpack(torch.stack(out).numpy())
    ok = len(frames) == B
    for f, fr in enumerate(frames):
        k = f % 4
        ok = ok and fr["info"].tolist() == [k, 0, 10 * f, f] and fr["model"].tolist() == [100 + f] * k
        ok = ok and np.allclose(fr["score"], f) and fr["pose"].shape == (k, 7) and np.allclose(fr["pose"][:, 0], f + np.arange(k) / 10)
    q.put((rank, bool(ok), (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_frame_partition_and_result_gather():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_frames, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert sorted(r[2] for r in res) == [(0, 3), (3, 6)]


def test_frame_range_rejects_uneven_split():
    from moped_b200.sharding import frame_range
    with pytest.raises(ValueError):
        frame_range(7, 2, 0)
    assert [frame_range(8, 4, r) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]


def _ransac_rank(rank, world, port, q):
    import torch.distributed as dist
    from moped_b200 import synth
    from moped_b200.sharding import cluster_partition
    from oracle import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cl = synth.make_ransac_clusters(5, 40, 0.3, seed=3)
    hy = synth.make_hypotheses(cl, 6, 5, seed=3)
    mine = cluster_partition(hy["hyp_cluster"], world, rank)
    cams = oracle.cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    n_in = np.full(len(hy["hyp_cluster"]), -7, np.int32)
    for h in mine:                                       # the per-rank work: its clusters' hypotheses, nothing else
        c = hy["hyp_cluster"][h]
        s = slice(cl["offsets"][c], cl["offsets"][c + 1])
        n_in[h] = oracle.hypothesis(cl["xy"][s], cl["xyz"][s], cl["image"][s], cams, hy["sample_pos"][h], hy["init_quat"][h], 200, 10.0, 6)[0]
    import torch
    t = torch.from_numpy(n_in)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)             # test-only gather of the disjoint results
    if rank == 0:
        q.put((mine.tolist(), t.numpy().tolist()))
    dist.destroy_process_group()


def test_two_rank_cluster_partition_covers_every_hypothesis_once():
    """RANSAC-heavy configuration over 2 ranks (gloo): every hypothesis is evaluated by exactly one rank (the owner of
    its cluster) and the union equals the single-process evaluation."""
    import torch.multiprocessing as mp
    from moped_b200 import synth
    from moped_b200.sharding import cluster_partition
    from oracle import oracle
    cl = synth.make_ransac_clusters(5, 40, 0.3, seed=3)
    hy = synth.make_hypotheses(cl, 6, 5, seed=3)
    parts = [cluster_partition(hy["hyp_cluster"], 2, r) for r in range(2)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(30))
    assert set(hy["hyp_cluster"][parts[0]]) == {0, 2, 4} and set(hy["hyp_cluster"][parts[1]]) == {1, 3}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_ransac_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    mine0, merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert mine0 == parts[0].tolist()
    cams = oracle.cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    single = []
    for h in range(30):
        c = hy["hyp_cluster"][h]
        s = slice(cl["offsets"][c], cl["offsets"][c + 1])
        single.append(oracle.hypothesis(cl["xy"][s], cl["xyz"][s], cl["image"][s], cams, hy["sample_pos"][h], hy["init_quat"][h], 200, 10.0, 6)[0])
    assert merged == single


# ---- feature extraction: frames of a batch are independent -> partitioned, no data-path collective ----
def _sift_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from moped_b200 import synth
    from moped_b200.sharding import frame_range
    from oracle import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_frames = 4
    frames = [synth.make_image(40 + i, 96, 128, n_blobs=120) for i in range(n_frames)]
    lo, hi = frame_range(n_frames, world, rank)
    mine = [oracle.sift(frames[f], True) for f in range(lo, hi)]               # (the oracle stands in for the GPU here)
    counts = torch.tensor([len(m[0]) for m in mine], dtype=torch.int32)
    digest = torch.tensor([float(np.float64(m[2]).sum()) for m in mine], dtype=torch.float64)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    all_digest = [torch.empty_like(digest) for _ in range(world)]
    dist.all_gather(all_counts, counts)                                          # bookkeeping only: the descriptors stay where they are
    dist.all_gather(all_digest, digest)
    single = [oracle.sift(f, True) for f in frames]
    ok_counts = torch.cat(all_counts).tolist() == [len(s[0]) for s in single]
    ok_digest = np.array_equal(torch.cat(all_digest).numpy(), np.array([float(np.float64(s[2]).sum()) for s in single]))
    q.put((rank, ok_counts, ok_digest, (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_frame_partition_of_feature_extraction():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sift_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[3] for r in res] == [(0, 2), (2, 4)]                               # contiguous blocks in frame order
    assert all(r[1] and r[2] for r in res), res


# ---- full RANSAC by cluster (moped3d's depth-aware POSE stage): contiguous cluster blocks + one seed offset ----
def _depth_ransac_rank(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from moped_b200.sharding import RANSAC_STREAM_STRIDE, ransac_cluster_range, ransac_shard_seed
    from oracle import oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle3d_pose import ALPHA, CAM, K, make_cluster
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    clusters = [make_cluster(700 + i, n=30, outliers=0.3) for i in range(5)]
    max_obj, seed, params = 2, 17, (12, 100, 5, 6, 8.0)
    cams = oracle.cameras(K[None], CAM[None])
    lo, hi = ransac_cluster_range(len(clusters), world, rank)
    shard_seed = ransac_shard_seed(seed, lo, max_obj)
    out = np.zeros((len(clusters) * max_obj, 9), np.float32)           # found, tests, pose: zero outside this rank's block
    for t in range((hi - lo) * max_obj):                               # what mc_pose_depth_ransac does with the shard's clusters and seed
        task_seed = (shard_seed + RANSAC_STREAM_STRIDE * (t + 1)) & 0xFFFFFFFFFFFFFFFF
        f, p, it = oracle.ransac_depth(clusters[lo + t // max_obj], cams, ALPHA, params, task_seed)
        out[lo * max_obj + t] = [f, it, *p]
    g = torch.from_numpy(out.view(np.int32).copy())
    dist.all_reduce(g, op=dist.ReduceOp.SUM)                            # disjoint blocks: the sum of bit patterns is the gather
    if rank == 0:
        q.put(((lo, hi), g.numpy().view(np.float32).tolist()))
    dist.destroy_process_group()


def test_two_rank_depth_ransac_by_cluster_is_independent_of_the_number_of_ranks():
    import torch.multiprocessing as mp
    from moped_b200.sharding import RANSAC_STREAM_STRIDE, ransac_cluster_range
    from oracle import oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle3d_pose import ALPHA, CAM, K, make_cluster
    assert [ransac_cluster_range(5, 2, r) for r in range(2)] == [(0, 3), (3, 5)]
    assert [ransac_cluster_range(7, 4, r) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 7)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29870 + os.getpid() % 100
    procs = [ctx.Process(target=_depth_ransac_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    (lo, hi), merged = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (lo, hi) == (0, 3)
    clusters = [make_cluster(700 + i, n=30, outliers=0.3) for i in range(5)]
    cams = oracle.cameras(K[None], CAM[None])
    found = 0
    for t in range(10):                                                 # the single-rank call: seed 17, global task index
        f, p, it = oracle.ransac_depth(clusters[t // 2], cams, ALPHA, (12, 100, 5, 6, 8.0), (17 + RANSAC_STREAM_STRIDE * (t + 1)) & 0xFFFFFFFFFFFFFFFF)
        assert merged[t][0] == float(f) and merged[t][1] == float(it), t
        if f:
            assert np.array_equal(np.array(merged[t][2:], np.float32), p), t
        found += f
    assert found >= 5


def _cluster_partition_frame_rank(rank, world, port, q):
    """mc_process_frame_sharded_dev's protocol with the oracle in the kernels' place: CLUSTER / FILTER replicated, the (cluster, try)
    tasks of POSE and POSE2 dealt by cluster, the task records all-gathered (gloo) and every task taken from its owner."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_chain as oc
    from conftest import cluster_points
    from oracle import oracle
    from moped_b200 import synth
    from moped_b200.sharding import ransac_task_owner, select_task_records
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    db = synth.make_db(6, 300, seed=4)
    fr = synth.make_frame(db, 500, n_visible=3, frame_id=2)
    dbn, qn = oracle.norm_rows(db["desc"]), oracle.norm_rows(fr["desc"])
    cams = oracle.cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    m, _, _ = oracle.match(dbn, db["xyz"], db["model_of_row"], 6, qn, fr["xy"], fr["image_idx"], 0.8)

    def pose_step_sharded(clusters, params, seed, objects):
        xy, xyz, img, tie, co = cluster_points(m, clusters)
        n_tasks = len(clusters["model"]) * params[2]
        rec = np.zeros((n_tasks, 8), np.float32)                       # found | pose[7]
        mine = 0
        for task in range(n_tasks):
            if ransac_task_owner(task, params[2], world) != rank:
                continue
            mine += 1
            k = task // params[2]
            sl = slice(co[k], co[k + 1])
            f, p, _ = oracle.ransac(xy[sl], xyz[sl], img[sl], tie[sl], cams, params, (seed + oc.GOLDEN * (task + 1)) & oc.M64)
            rec[task, 0] = float(f)
            rec[task, 1:] = p
        out = [torch.empty((n_tasks, 8), dtype=torch.float32) for _ in range(world)]
        dist.all_gather(out, torch.from_numpy(rec))
        sel = select_task_records(torch.stack(out).numpy(), params[2])
        for task in range(n_tasks):
            if sel[task, 0]:
                objects.append((int(clusters["model"][task // params[2]]), sel[task, 1:].copy()))
        return mine, n_tasks

    cl = oracle.cluster(m, 1)
    objects = []
    mine1, n1 = pose_step_sharded(cl, oc.POSE1, 1, objects)
    objects, cl2, _ = oc.filter_step(m, cams, objects, oc.FILTER1)
    mine2, n2 = pose_step_sharded(cl2, oc.POSE2, 2, objects)
    objects, _, score = oc.filter_step(m, cams, objects, oc.FILTER2)
    want = oc.frame(dbn, db["xyz"], db["model_of_row"], 6, qn, fr["xy"], fr["image_idx"], synth.K_DEFAULT, synth.CAM_IDENTITY)
    ok = len(objects) == len(want["model"]) and len(objects) >= 2
    ok = ok and [o[0] for o in objects] == want["model"].tolist()
    ok = ok and all(np.array_equal(o[1].astype(np.float32), w) for o, w in zip(objects, want["pose"])) and np.array_equal(score, want["score"])
    q.put((rank, bool(ok), (mine1, n1, mine2, n2), len(objects)))
    dist.destroy_process_group()


def test_two_rank_frame_with_ransac_distributed_by_cluster_equals_the_single_process_chain():
    """north_star: "RANSAC work is distributed by cluster" inside the frame pipeline. Bit-identical objects on both ranks, every task run
    by exactly one rank."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cluster_partition_frame_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    (a1, n1, a2, n2), (b1, _, b2, _) = res[0][2], res[1][2]
    assert a1 + b1 == n1 and a2 + b2 == n2 and n1 > 0, res
    assert res[0][3] == res[1][3]
