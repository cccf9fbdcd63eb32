"""moped3d's FILTER_PROJECTION_DEPTH_CPU (filter/FILTER_PROJECTION_DEPTH_CPU.hpp): the C restatement (oracle/moped_oracle.c: mo_filter_depth,
mo_filter_depth_select) against the class compiled unmodified from /root/reference (oracle/ref3d_harness.cpp: ref3d_filter_depth) —
keep flags, scores (projection score minus the depth penalty), rebuilt clusters and the random choice of test points identical to the
strict-IEEE build bit for bit; the build with the reference's own -ffast-math flags agrees on every decision and to 1e-4 on the scores."""
import numpy as np
import pytest

from moped_b200 import synth
from oracle import oracle, ref3d

pytestmark = pytest.mark.skipif(not ref3d.available(), reason="oracle/_ref/libmoped3d_ref*.so not built (needs /root/reference at build time)")

PARAMS = [(5, 4096.0, 16384.0, 2.0, 0.05, 0.2), (4, 1024.0, 4096.0, 1.0, 0.02, 0.0), (6, 4096.0, 4096.0, 3.0, 0.1, 0.9)]


def run_both(sc, params, sample_size, seed, strict):
    ref3d.use_strict(strict)
    try:
        r = ref3d.filter_depth(np.diff(sc["model_offsets"]), sc["model_xyz"], sc["match_offsets"], sc["match_xy"], sc["match_xyz"], sc["obj_model"],
                               sc["obj_pose"], params, sample_size, seed, sc["K"], sc["cam_pose"], sc["depth_K"], sc["depth_pose"], sc["depth"], sc["fill"])
    finally:
        ref3d.use_strict(False)
    cams = oracle.cameras(sc["K"], sc["cam_pose"])
    dcam = oracle.cameras(sc["depth_K"], sc["depth_pose"])
    to, txyz = oracle.filter_depth_test_points(sc["model_offsets"], sc["model_xyz"], sample_size, seed)
    m = dict(offsets=sc["match_offsets"], image=sc["match_image"], xy=sc["match_xy"], xyz=sc["match_xyz"])
    o = oracle.filter_depth(m, cams, sc["obj_model"], sc["obj_pose"], params, to, txyz, dcam, sc["depth"], sc["fill"])
    return r, o


@pytest.mark.parametrize("params", PARAMS)
@pytest.mark.parametrize("sample_size", [100000, 40])
def test_bit_identical_to_the_strict_build(params, sample_size):
    checked = pruned = penalised = 0
    for seed in range(4):
        sc = synth.make_filter_depth_scene(seed)
        r, o = run_both(sc, params, sample_size, 1000 + seed, strict=True)
        assert np.array_equal(r["keep"], o["keep"]), (seed, r["keep"], o["keep"])
        assert np.array_equal(r["score"], o["score"]), (seed, r["score"], o["score"])
        assert np.array_equal(r["offsets"], o["offsets"]) and np.array_equal(r["members"], o["members"]), seed
        checked += len(o["keep"]); pruned += int((~o["keep"]).sum())
        # the penalty is at work: the same frame through the plain projection filter scores some objects higher
        cams = oracle.cameras(sc["K"], sc["cam_pose"])
        m = dict(offsets=sc["match_offsets"], image=sc["match_image"], xy=sc["match_xy"], xyz=sc["match_xyz"])
        plain = oracle.filter_objects(m, cams, sc["obj_model"], sc["obj_pose"], (params[0], params[1], params[3]))
        penalised += int((plain["score"] > o["score"]).sum())
    assert checked >= 40 and pruned >= 8
    if params[5] < 0.9:
        assert penalised >= 4


def test_fast_math_build_takes_the_same_decisions():
    for seed in range(3):
        sc = synth.make_filter_depth_scene(seed)
        r, o = run_both(sc, PARAMS[0], 100000, 7, strict=False)
        assert np.array_equal(r["keep"], o["keep"])
        assert np.allclose(r["score"], o["score"], rtol=1e-4, atol=1e-4)
        assert np.array_equal(r["offsets"], o["offsets"]) and np.array_equal(r["members"], o["members"])


def test_empty_object_list_and_model_without_matches():
    sc = synth.make_filter_depth_scene(2)
    sc = dict(sc)
    sc["obj_model"] = sc["obj_model"][:0]; sc["obj_pose"] = sc["obj_pose"][:0]
    r, o = run_both(sc, PARAMS[0], 100000, 3, strict=True)
    assert len(o["keep"]) == 0 and len(r["keep"]) == 0 and len(o["offsets"]) == 1
    sc = dict(synth.make_filter_depth_scene(3))
    lo, hi = sc["match_offsets"][1], sc["match_offsets"][2]                   # model 1 loses all its matches
    keep = np.r_[np.arange(lo), np.arange(hi, sc["match_offsets"][-1])]
    sc["match_xy"], sc["match_xyz"], sc["match_image"] = sc["match_xy"][keep], sc["match_xyz"][keep], sc["match_image"][keep]
    mo = sc["match_offsets"].copy(); mo[2:] -= hi - lo; sc["match_offsets"] = mo
    r, o = run_both(sc, PARAMS[0], 100000, 3, strict=True)
    assert np.array_equal(r["keep"], o["keep"]) and np.array_equal(r["score"], o["score"])
    assert not o["keep"][sc["obj_model"] == 1].any()
