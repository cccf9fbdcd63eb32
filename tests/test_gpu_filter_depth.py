"""moped3d's FILTER_PROJECTION_DEPTH on the GPU (mc_filter_projection_depth, filter.cu: k_filter_depth_adjust) against the oracle
(oracle/moped_oracle.c: mo_filter_depth, pinned bit for bit to the strict-IEEE build of the reference's class by
tests/test_oracle3d_filter.py): keep flags, scores (projection score minus depth penalty, bit for bit), rebuilt clusters — with all
keypoints as test points and with a drawn sample; and the stage class inside moped3d's own MopedPipeline next to the CPU class."""
import os
import re
import subprocess

import numpy as np
import pytest

from moped_b200 import synth
from oracle import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARAMS = [(5, 4096.0, 16384.0, 2.0, 0.05, 0.2), (4, 1024.0, 4096.0, 1.0, 0.02, 0.0), (6, 4096.0, 4096.0, 3.0, 0.1, 0.9)]


@pytest.mark.parametrize("params", PARAMS)
@pytest.mark.parametrize("sample_size", [100000, 40])
def test_filter_depth_equals_the_oracle_bit_for_bit(gpu_ctx, params, sample_size):
    pruned = kept = 0
    for seed in range(4):
        sc = synth.make_filter_depth_scene(seed)
        cams = oracle.cameras(sc["K"], sc["cam_pose"])
        dcam = oracle.cameras(sc["depth_K"], sc["depth_pose"])
        to, txyz = oracle.filter_depth_test_points(sc["model_offsets"], sc["model_xyz"], sample_size, 99 + seed)
        m = dict(offsets=sc["match_offsets"], image=sc["match_image"], xy=sc["match_xy"], xyz=sc["match_xyz"])
        want = oracle.filter_depth(m, cams, sc["obj_model"], sc["obj_pose"], params, to, txyz, dcam, sc["depth"], sc["fill"])
        gpu_ctx.set_cameras(sc["K"], sc["cam_pose"])
        got = gpu_ctx.filter_depth(m, sc["obj_model"], sc["obj_pose"], params, to, txyz, sc["depth_K"], sc["depth_pose"], sc["depth"], sc["fill"])
        assert np.array_equal(got["keep"], want["keep"]), (seed, got["keep"], want["keep"])
        assert np.array_equal(got["score"], want["score"]), (seed, got["score"], want["score"])
        assert np.array_equal(got["offsets"], want["offsets"]) and np.array_equal(got["members"], want["members"]), seed
        pruned += int((~want["keep"]).sum()); kept += int(want["keep"].sum())
    assert pruned >= 8 and (kept >= 3 or params[0] == 6)


def test_filter_depth_edge_cases(gpu_ctx):
    sc = synth.make_filter_depth_scene(5)
    m = dict(offsets=sc["match_offsets"], image=sc["match_image"], xy=sc["match_xy"], xyz=sc["match_xyz"])
    to, txyz = oracle.filter_depth_test_points(sc["model_offsets"], sc["model_xyz"], 100000, 1)
    gpu_ctx.set_cameras(sc["K"], sc["cam_pose"])
    p = PARAMS[0]
    # no objects
    got = gpu_ctx.filter_depth(m, sc["obj_model"][:0], sc["obj_pose"][:0], p, to, txyz, sc["depth_K"], sc["depth_pose"], sc["depth"], sc["fill"])
    assert len(got["keep"]) == 0 and got["offsets"].tolist() == [0]
    # no test points at all: the penalty is 0 and the result is the plain projection filter's
    zero = np.zeros(len(sc["model_offsets"]), np.int32)
    got = gpu_ctx.filter_depth(m, sc["obj_model"], sc["obj_pose"], p, zero, np.zeros((0, 3), np.float32), sc["depth_K"], sc["depth_pose"], sc["depth"], sc["fill"])
    plain = gpu_ctx.filter(m, sc["obj_model"], sc["obj_pose"], (p[0], p[1], p[3]))
    assert np.array_equal(got["keep"], plain["keep"]) and np.array_equal(got["score"], plain["score"])
    assert np.array_equal(got["offsets"], plain["offsets"]) and np.array_equal(got["members"], plain["members"])
    # a depth map full of holes (fill distance > 0 everywhere): no usable point, penalty 0
    got = gpu_ctx.filter_depth(m, sc["obj_model"], sc["obj_pose"], p, to, txyz, sc["depth_K"], sc["depth_pose"], sc["depth"], np.ones_like(sc["fill"]))
    assert np.array_equal(got["score"], plain["score"])
    # poses that put test points behind / on the camera plane (division by zero, NaN coordinates) follow the oracle
    pose = sc["obj_pose"].copy(); pose[:, 6] = 0.0
    cams = oracle.cameras(sc["K"], sc["cam_pose"]); dcam = oracle.cameras(sc["depth_K"], sc["depth_pose"])
    want = oracle.filter_depth(m, cams, sc["obj_model"], pose, p, to, txyz, dcam, sc["depth"], sc["fill"])
    got = gpu_ctx.filter_depth(m, sc["obj_model"], pose, p, to, txyz, sc["depth_K"], sc["depth_pose"], sc["depth"], sc["fill"])
    assert np.array_equal(got["keep"], want["keep"]) and np.array_equal(got["score"], want["score"])


def test_dropin_filter_projection_depth_inside_moped3d_api(tmp_path):
    """FILTER_PROJECTION_DEPTH_CPU and FILTER_PROJECTION_DEPTH_CUDA registered in two moped3d MopedPipelines (oracle/ref3d_filter_dropin.cpp),
    identical FrameData and srand(): the same surviving objects in the same order with the same score bits, the same clusters — with all
    keypoints as test points and with 40 drawn ones (the stage class draws them with rand() like the reference)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "moped3d_filter_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/moped3d_filter_dropin not built (needs /root/reference at build time)")
    for seed, sample in ((11, 100000), (12, 40)):
        sc = synth.make_filter_depth_scene(seed)
        path = str(tmp_path / f"filter_case_{seed}.bin")
        nm = len(sc["model_offsets"]) - 1
        with open(path, "wb") as f:
            np.array([sc["width"], sc["height"], nm, len(sc["obj_model"]), sample], np.int32).tofile(f)
            sc["K"].astype(np.float32).tofile(f)
            np.diff(sc["model_offsets"]).astype(np.int32).tofile(f)
            np.diff(sc["match_offsets"]).astype(np.int32).tofile(f)
            sc["depth"].astype(np.float32).tofile(f); sc["fill"].astype(np.float32).tofile(f)
            sc["model_xyz"].astype(np.float32).tofile(f)
            np.concatenate([sc["match_xy"], sc["match_xyz"]], axis=1).astype(np.float32).tofile(f)
            for o in range(len(sc["obj_model"])):
                np.array([sc["obj_model"][o]], np.int32).tofile(f)
                sc["obj_pose"][o].astype(np.float32).tofile(f)
        r = subprocess.run([exe, path], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert "CONFIG FILTER:0:FILTER_PROJECTION_DEPTH_CUDA/PlausibleSqDistance=16384" in r.stdout, r.stdout
        mt = re.search(r"STEP FILTER same_objects=(\d) cpu_objects=(\d+) gpu_objects=(\d+) same_clusters=(\d) input_objects=(\d+)", r.stdout)
        assert mt, r.stdout
        assert mt.group(1) == "1" and mt.group(4) == "1" and 1 <= int(mt.group(2)) < int(mt.group(5)), r.stdout
