"""Host-side sharding of the model database by object: contiguous object ranges balanced by
descriptor count. Rank r uploads rows [row_lo, row_hi) with row_base = row_lo (mc_db_upload) plus the two small
global tables (mc_db_set_global_tables); per-query (row, distance) pairs of all shards are exchanged with one
all-gather and merged by mc_match_merge_dev (smaller distance first, then smaller global row id)."""
from __future__ import annotations

import numpy as np


def shard_objects(n_pts: np.ndarray, world: int):
    """-> list of (obj_lo, obj_hi, row_lo, row_hi) per rank."""
    n_pts = np.asarray(n_pts)
    cum = np.concatenate([[0], np.cumsum(n_pts)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        b = int(np.searchsorted(cum, total * r / world))
        bounds.append(min(max(b, bounds[-1]), len(n_pts)))
    bounds.append(len(n_pts))
    return [(bounds[r], bounds[r + 1], int(cum[bounds[r]]), int(cum[bounds[r + 1]])) for r in range(world)]


def frame_range(n_frames: int, world: int, rank: int):
    """Frames of a batch that rank `rank` runs through CLUSTER..FILTER2 after the merged MATCH: contiguous blocks
    of n_frames / world (so that the all-gathered per-rank result blocks are in frame order)."""
    if n_frames % world:
        raise ValueError(f"{n_frames} frames do not split evenly over {world} ranks")
    per = n_frames // world
    return rank * per, (rank + 1) * per


class ResultBlock:
    """Layout of one rank's result block for `frames` frames with `max_objects` slots each, in int32 words:
    info[frames,4] | model[frames,MO] | score[frames,MO] (f32 bits) | pose[frames,MO*7] (f32 bits).
    mc_process_frames_matched_dev writes the four regions in place; one all-gather moves the block."""

    def __init__(self, frames: int, max_objects: int):
        self.frames, self.mo = frames, max_objects
        self.o_info = 0
        self.o_model = frames * 4
        self.o_score = frames * (4 + max_objects)
        self.o_pose = frames * (4 + 2 * max_objects)
        self.words = frames * (4 + 9 * max_objects)

    def unpack(self, blocks: np.ndarray):
        """blocks: [world, words] int32 (the all-gathered blocks) -> list of per-frame dict(model, pose, score, info)
        in frame order."""
        blocks = np.asarray(blocks, dtype=np.int32).reshape(-1, self.words)
        out = []
        for b in blocks:
            info = b[self.o_info:self.o_model].reshape(self.frames, 4)
            model = b[self.o_model:self.o_score].reshape(self.frames, self.mo)
            score = b[self.o_score:self.o_pose].view(np.float32).reshape(self.frames, self.mo)
            pose = b[self.o_pose:self.words].view(np.float32).reshape(self.frames, self.mo, 7)
            for f in range(self.frames):
                k = min(int(info[f, 0]), self.mo)
                out.append(dict(model=model[f, :k].copy(), pose=pose[f, :k].copy(), score=score[f, :k].copy(), info=info[f].copy()))
        return out


def cluster_partition(hyp_cluster: np.ndarray, world: int, rank: int) -> np.ndarray:
    """RANSAC work distributed by cluster (north_star; BASELINE configs[3]): cluster c belongs to rank c % world, so a
    rank owns every hypothesis of its clusters and no data-path collective is needed. Returns the indices (ascending)
    of the hypotheses rank `rank` evaluates; the union over ranks is a partition of range(len(hyp_cluster))."""
    return np.nonzero(np.asarray(hyp_cluster) % world == rank)[0]


RANSAC_STREAM_STRIDE = 0x9E3779B97F4A7C15      # task t of a call draws from seed + stride * (t + 1)  (pose.cu / pose_depth.cu)


def ransac_cluster_range(n_clusters: int, world: int, rank: int):
    """Full RANSAC (mc_pose_ransac / mc_pose_depth_ransac) distributed by cluster: rank r runs the contiguous cluster block
    [lo, hi) — contiguous, so that one seed offset (ransac_shard_seed) keeps every task on the random stream it has in a
    single-GPU call and the result does not depend on the number of ranks."""
    per, extra = divmod(n_clusters, world)
    lo = rank * per + min(rank, extra)
    return lo, lo + per + (1 if rank < extra else 0)


def ransac_shard_seed(seed: int, first_cluster: int, max_objects_per_cluster: int) -> int:
    """The `seed` a shard passes so that its LOCAL task t (cluster first_cluster + t // MaxObjectsPerCluster) draws from the
    stream of GLOBAL task first_cluster * MaxObjectsPerCluster + t: seed + stride * first_task, modulo 2^64."""
    return (seed + RANSAC_STREAM_STRIDE * first_cluster * max_objects_per_cluster) & 0xFFFFFFFFFFFFFFFF


def ransac_task_owner(task: int, max_objects_per_cluster: int, world: int) -> int:
    """Frame pipeline with the RANSAC work distributed by cluster (mc_process_frame_sharded_dev): task t = (cluster, try) belongs to the
    rank that owns its cluster, cluster c -> rank c % world (the cluster count is only known on the device, so the deal is round-robin)."""
    return (task // max_objects_per_cluster) % world


def select_task_records(records: np.ndarray, max_objects_per_cluster: int) -> np.ndarray:
    """What k_shard_select does after the all-gather: records[r][t] is rank r's record of task t (zeros unless r owns t); the result
    takes every task from its owner."""
    world, n_tasks = records.shape[:2]
    owner = (np.arange(n_tasks) // max_objects_per_cluster) % world
    return records[owner, np.arange(n_tasks)]
