"""Writer of moped3d's benchmark log — the text format of MopedBench (moped3d/libmoped/src/MopedBench.cpp:60-227, Benchmark.hpp) —
from the flat arrays the C ABI returns, so that runs of the CUDA stages can be diffed line by line against logs of the reference
("the MopedBench dump format as a parity-log format"). Pure host code (numpy fp32 in the reference's expression
order for the projected hulls); checked against the compiled MopedBench itself by tests/test_bench_log.py.

Lines, in the order MopedBench::init wires them (MopedBench.cpp:211-222):
    PRE:CLUSTER:MATCH ModelNum:m;Idx:i;2DLoc:[x y];DepthValid:b;XYZ: [X Y Z] ;Depth:d;FillDistance:f      (one per match)
    TIME:<step>:<seconds>
    POST:CLUSTER:CLUSTERCOUNT:a|b|...          PRE:/POST:FILTER:CLUSTERCOUNT:...          PRE:POSE:CLUSTERCOUNT:...
    POST:POSE:OBJHULL:<model>:x;y:x;y...       POST:FILTER2:OBJHULL:...                   (convex hull of the projected model points)
    OBJ: <model> [tx ty tz] [qx qy qz qw] <score>                                           (allDone)
"""
from __future__ import annotations

import numpy as np

F = np.float32


def _g(v) -> str:
    """operator<<(ostream&, float): %g with 6 significant digits."""
    return "%g" % float(v)


def _pt(p) -> str:
    return "[" + " ".join(_g(x) for x in p) + "]"


def _tm(q, t):
    """TransformMatrix::init (moped.hpp:181-188), fp32, the reference's expression order."""
    q = [F(x) for x in q]
    two = F(2)
    one = F(1)
    R = [[one - two * q[1] * q[1] - two * q[2] * q[2], two * q[0] * q[1] - two * q[3] * q[2], two * q[0] * q[2] + two * q[3] * q[1]],
         [two * q[0] * q[1] + two * q[3] * q[2], one - two * q[0] * q[0] - two * q[2] * q[2], two * q[1] * q[2] - two * q[3] * q[0]],
         [two * q[0] * q[2] - two * q[3] * q[1], two * q[1] * q[2] + two * q[3] * q[0], one - two * q[0] * q[0] - two * q[1] * q[1]]]
    return R, [F(x) for x in t]


def project(pose7, xyz, K4, cam_pose7):
    """project() (moped.hpp:318-344): model point -> world (object pose) -> camera (inverse of the camera pose) -> pixels."""
    R, t = _tm(pose7[:4], pose7[4:])
    C, ct = _tm(cam_pose7[:4], cam_pose7[4:])
    o = [F(v) for v in xyz]
    p = [o[0] * R[r][0] + o[1] * R[r][1] + o[2] * R[r][2] + t[r] for r in range(3)]
    d = [p[r] - ct[r] for r in range(3)]
    c = [d[0] * C[0][k] + d[1] * C[1][k] + d[2] * C[2][k] for k in range(3)]
    if c[2] < 0.001:
        return (F(np.finfo(np.float32).max), F(np.finfo(np.float32).max))
    K4 = [F(v) for v in K4]
    return (c[0] / c[2] * K4[0] + K4[2], c[1] / c[2] * K4[1] + K4[3])


def convex_hull(points):
    """getConvexHull (moped.hpp:346-380): sort, then one sweep forward and one back over a list, popping while det <= 0."""
    pts = sorted(points)                          # Pt::operator< is lexicographic
    half, hull = [pts[0]], []                     # lists with the FRONT at index 0
    p, direction = 1, 1
    while p != -1 and p < len(pts):
        half.insert(0, pts[p])
        convex = False
        while not convex and len(half) > 2:
            p2, p1, p0 = half[0], half[1], half[2]
            det = ((p0[0] - p1[0]) * (p2[1] - p1[1])) - ((p2[0] - p1[0]) * (p0[1] - p1[1]))
            if det <= 0:
                del half[1]
            else:
                convex = True
        if p == len(pts) - 1:
            half.pop(0)
            hull = half + hull
            half = [pts[p]]
            direction = -1
        p += direction
    half.pop(0)
    return half + hull


class MopedBenchLog:
    """Collects the lines MopedBench would write for one frame. `frame` is a dict of flat arrays:
    n_matches[m]; match_xy, match_world (n x 3), match_depth, match_fill, match_valid, match_image (concatenated in model order);
    cluster_model[c], cluster_offsets[c+1], cluster_members; obj_model[o], obj_pose[o,7] (quaternion xyzw + translation), obj_score[o];
    model_names[m], model_xyz (list of [n_m,3] arrays); K (4), cam_pose (7)."""

    def __init__(self):
        self.lines = []

    def _cluster_count(self, tag, fr):
        counts = np.bincount(np.asarray(fr["cluster_model"], dtype=np.int64), minlength=len(fr["n_matches"])) if len(fr["cluster_model"]) else \
            np.zeros(len(fr["n_matches"]), np.int64)
        self.lines.append(f"{tag}:CLUSTERCOUNT:" + "|".join(str(int(c)) for c in counts))

    def _matches(self, tag, fr):
        k = 0
        for m, n in enumerate(fr["n_matches"]):
            for _ in range(int(n)):
                w = fr["match_world"][k]
                self.lines.append(f"{tag}:MATCH ModelNum:{m};Idx:{int(fr['match_image'][k])};2DLoc:{_pt(fr['match_xy'][k])};"
                                  f"DepthValid:{int(bool(fr['match_valid'][k]))};XYZ: [{_g(w[0])} {_g(w[1])} {_g(w[2])}] ;"
                                  f"Depth:{_g(fr['match_depth'][k])};FillDistance:{_g(fr['match_fill'][k])}")
                k += 1

    def _hulls(self, tag, fr):
        for o in range(len(fr["obj_model"])):
            m = int(fr["obj_model"][o])
            pts = [project(fr["obj_pose"][o], x, fr["K"], fr["cam_pose"]) for x in fr["model_xyz"][m]]
            hull = convex_hull(pts)
            self.lines.append(f"{tag}:OBJHULL:{fr['model_names'][m]}" + "".join(f":{_g(p[0])};{_g(p[1])}" for p in hull))

    def step(self, name, fr, seconds):
        """beforeAlgorithm + afterAlgorithm of one step (MopedBench.cpp:185-199) with the actions of init()."""
        pre = {"CLUSTER": self._matches, "FILTER": self._cluster_count, "POSE": self._cluster_count}.get(name)
        post = {"CLUSTER": self._cluster_count, "FILTER": self._cluster_count, "POSE": self._hulls, "FILTER2": self._hulls}.get(name)
        if pre:
            pre("PRE:" + name, fr)
        self.lines.append(f"TIME:{name}:{_g(seconds)}")
        if post:
            post("POST:" + name, fr)

    def all_done(self, fr):
        for o in range(len(fr["obj_model"])):
            p = fr["obj_pose"][o]
            self.lines.append(f"OBJ: {fr['model_names'][int(fr['obj_model'][o])]} {_pt(p[4:7])} {_pt(p[0:4])} {_g(fr['obj_score'][o])}")

    def text(self) -> str:
        return "\n".join(self.lines) + "\n"
