"""moped3d's MATCH step: the host logic of MATCH_ADAPTIVE_CUDA (moped_b200/stages/MATCH_ADAPTIVE_CUDA.hpp + adaptive_ratio.hpp: per-model
control points from bounding box / intrinsics / feature count, the depth- and fill-distance-dependent ratio threshold, the depth
cut, in-place normalisation, match assembly) against the reference's own MATCH_ADAPTIVE_FLANN_CPU, compiled UNMODIFIED inside
moped3d's tree over a stand-in cv::flann::Index that searches exhaustively (oracle/ref3d_match_dropin.cpp — OpenCV's FLANN is
external, unpinned and absent here). Without a GPU the CUDA class's nearest-neighbour step is driven by the same
exhaustive search ("host" mode); with one, tests/test_gpu_depth_pose.py runs the class as shipped (mc_match)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "moped3d_match_dropin")


def run_dropin(seed, mode=None):
    r = subprocess.run([EXE, str(seed)] + ([mode] if mode else []), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    step = [l for l in r.stdout.splitlines() if l.startswith("STEP MATCH")][0]
    return dict(kv.split("=") for kv in step.split()[2:]), r.stdout


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/moped3d_match_dropin not built (needs /root/reference at build time)")
def test_adaptive_ratio_logic_equals_the_reference_class():
    for seed in (1, 2, 3, 4):
        info, out = run_dropin(seed, "host")
        assert info["same"] == "1" and info["normalised_features_same"] == "1" and info["normalised_models_same"] == "1", (seed, info)
        assert 200 < int(info["matches"]) < 600, info            # the adaptive threshold both accepts and rejects planted copies
        per_model = {}
        for line in out.splitlines():
            if line.startswith("MATCH cpu"):
                per_model[line.split()[2]] = per_model.get(line.split()[2], 0) + 1
        assert len(per_model) == 5, per_model                    # sparse and dense models (both sides of the density sigmoid) all match
    for key in ("MinRatioMin", "MinRatioMax", "MaxRatioMin", "MaxRatioMax", "DimensionPeak", "DimensionFade", "NumTrees", "DescriptorType",
                "DescriptorSize"):
        assert "CONFIG MATCH_SIFT:0:MATCH_ADAPTIVE_CUDA/%s=" % key in out, key


def test_moped3d_cuda_pipeline_registers_with_the_reference_config_keys(tmp_path):
    """moped_b200/stages/pipeline3d_cuda.hpp inside moped3d's own MopedPipeline (no device needed to register and read the
    configuration): six algorithms for the six recognition steps of config.hpp:41-49, each exposing the config keys of the CPU
    class it replaces (9 / 6 / 5 / 3 / 5 / 3)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "moped3d_pose_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/moped3d_pose_dropin not built (needs /root/reference at build time)")
    case = tmp_path / "empty.bin"
    import numpy as np
    with open(case, "wb") as f:
        np.array([0], np.int32).tofile(f); np.zeros(4, np.float32).tofile(f)
    out = subprocess.run([exe, str(case), "3"], capture_output=True, text=True, timeout=60).stdout
    assert "STEP REGISTER algs=6" in out
    keys = {}
    for line in out.splitlines():
        if line.startswith("CONFIG "):
            step, _, rest = line[7:].split(":", 2)
            keys.setdefault((step, rest.split("/")[0]), []).append(rest.split("/")[1].split("=")[0])
    assert {k: len(v) for k, v in keys.items()} == {
        ("MATCH_SIFT", "MATCH_ADAPTIVE_CUDA"): 9, ("CLUSTER", "CLUSTER_LINKAGE_CUDA"): 6,
        ("POSE", "POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CUDA"): 5, ("FILTER", "FILTER_PROJECTION_CUDA"): 3,
        ("POSE2", "POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CUDA"): 5, ("FILTER2", "FILTER_PROJECTION_CUDA"): 3}
