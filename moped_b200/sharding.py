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
