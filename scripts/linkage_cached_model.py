"""Algorithmic model for the NEXT version of k_link_agglomerate (moped_b200/csrc/linkage.cu): the same merge sequence as
CLUSTER_LINKAGE_CPU::hierarchicalCluster — including its scan-order tie rule, the stale entry of the merged-away cluster and the
skipped list element — but with a cached maximum per row, so that a merge costs O(n) + the rows whose cached maximum was invalidated
instead of an O(n^2) scan of all live pairs (profiles/launches_r1k_linkage_summary.md: the scan is 98 % of the stage).

    python scripts/linkage_cached_model.py        # self-check against the oracle on tie-heavy and real-valued matrices

Invariants kept per live list position a (list L of cluster ids, ascending, with lazy erase of the merged-away id):
    best[a] = max over later positions b of D[L[a]][L[b]]; arg[a] = the SMALLEST such b          (strict > in scan order)
The pass maximum is the max over positions a not in {r, r+1} (r = position of the merged-away id, erased after the pass) of best[a],
ties to the smallest a. After a merge (tU absorbs rV): row/column tU of D changes -> best[] of tU's own row is recomputed, rows before
tU compare the new D[.][tU] against their cache (recompute only if their arg was tU and the value dropped, or on an equal value at a
smaller position); rV's column values do not change (stale by design) and rV leaves the list after the next pass -> rows whose arg
pointed at or beyond rV are repaired then. Only average linkage (the shipped configuration) is modelled; minimum / maximum linkage
recompute D from the member lists exactly as before and reuse the same cache maintenance.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def agglomerate_cached(K, cutoff, min_pts):
    K = np.asarray(K, np.float32)
    n = len(K)
    D = K.copy()
    for i in range(n):                      # distances[j*N+i] = distances[i*N+j] = K->getProb(i, j), j >= i
        for j in range(i, n):
            D[j, i] = D[i, j] = K[j, i]
    L = list(range(n))
    members = [[i] for i in range(n)]
    best = np.full(n, -1.0, np.float32)     # indexed by cluster id (row), not by position
    arg = np.full(n, -1, np.int64)          # cluster id of the arg-max column

    def recompute(pos):
        i = L[pos]
        b, a = np.float32(-1), -1
        for q in range(pos + 1, len(L)):
            v = D[i, L[q]]
            if v > b:
                b, a = v, L[q]
        best[i], arg[i] = b, a

    for pos in range(n):
        recompute(pos)
    remove = -1
    rescans = 0
    while True:
        r = L.index(remove) if remove in L else -1
        mx, p1, p2 = np.float32(-1), -1, -1
        for pos, i in enumerate(L):
            if r >= 0 and pos in (r, r + 1):
                continue
            if best[i] > mx:
                mx, p1, p2 = best[i], i, int(arg[i])      # the pair is fixed here: the stale column may be the winner
        if r >= 0:                          # the erase the reference does while scanning
            gone = L.pop(r)
            for pos in range(r):            # rows before it that pointed at it lose their maximum
                if arg[L[pos]] == gone:
                    recompute(pos); rescans += 1
        if mx < cutoff:
            break
        sU, sR = len(members[p1]), len(members[p2])
        while members[p2]:
            members[p1].append(members[p2].pop())
        remove = p2
        old_col = D[:, p1].copy()
        new = np.empty(n, np.float32)
        for i in range(n):
            new[i] = np.float32((1.0 / (sU + sR)) * float(np.float32(np.float32(sU) * D[p1, i]) + np.float32(np.float32(sR) * D[p2, i])))
        D[p1, :] = new
        D[:, p1] = new
        pos1 = L.index(p1)
        recompute(pos1); rescans += 1
        for pos in range(pos1):             # rows before tU: their entry in column tU changed
            i = L[pos]
            v = D[i, p1]
            if arg[i] == p1:
                if v < old_col[i]:
                    recompute(pos); rescans += 1
                else:
                    best[i] = v
            elif v > best[i] or (v == best[i] and L.index(int(arg[i])) > pos1):
                best[i], arg[i] = v, p1
    off, mem = [0], []
    for i in range(n):
        if len(members[i]) > min_pts:
            mem += members[i]
            off.append(len(mem))
    return np.array(off, np.int32), np.array(mem, np.int32), rescans


if __name__ == "__main__":
    from oracle import oracle
    rng = np.random.default_rng(0)
    checked = 0
    for case in range(300):
        n = int(rng.integers(2, 60))
        if case % 2:
            levels = int(rng.integers(2, 6))
            K = rng.integers(0, levels + 1, (n, n)).astype(np.float32) / levels
        else:
            K = rng.random((n, n)).astype(np.float32)
        K = np.maximum(K, K.T); np.fill_diagonal(K, 1.0)
        cutoff = float(rng.choice([0.2, 0.5, 0.75, 1.0]))
        minpts = int(rng.integers(0, 4))
        oo, om = oracle.linkage_agglomerate(K, cutoff, minpts, 1)
        co, cm, rescans = agglomerate_cached(K, cutoff, minpts)
        assert np.array_equal(oo, co) and np.array_equal(om, cm), (case, n, cutoff, minpts)
        checked += 1
    print(f"cached-maximum agglomeration == oracle on {checked} matrices (tie-heavy and real-valued)")
