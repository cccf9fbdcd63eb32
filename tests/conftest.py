import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "moped_golden.npz")))


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def ref_mod():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return ref


@pytest.fixture(scope="session")
def small_case(oracle_mod):
    """20 objects x 1000 descriptors, one 2000-feature frame with 4 planted objects; descriptors normalised."""
    from moped_b200 import synth
    db = synth.make_db(20, 1000)
    fr = synth.make_frame(db, 2000, n_visible=4)
    dbn = oracle_mod.norm_rows(db["desc"])
    qn = oracle_mod.norm_rows(fr["desc"])
    return dict(db=db, fr=fr, dbn=dbn, qn=qn, n_obj=20)


@pytest.fixture(scope="session")
def gpu_ctx():
    from moped_b200 import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()


def golden_matches(g):
    return dict(offsets=g["match_offsets"], image=g["match_image"], xy=g["match_xy"], xyz=g["match_xyz"])


def cluster_points(matches, clusters):
    """Flatten clusters into contiguous per-cluster point arrays (cluster CSR over points)."""
    xy, xyz, img, tie, co = [], [], [], [], [0]
    for c in range(len(clusters["model"])):
        mem = clusters["members"][clusters["offsets"][c]:clusters["offsets"][c + 1]]
        lo = matches["offsets"][clusters["model"][c]]
        xy.append(matches["xy"][lo + mem]); xyz.append(matches["xyz"][lo + mem]); img.append(matches["image"][lo + mem]); tie.append(mem)
        co.append(co[-1] + len(mem))
    if not xy:
        return np.zeros((0, 2), np.float32), np.zeros((0, 3), np.float32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.array(co, np.int32)
    return (np.concatenate(xy).astype(np.float32), np.concatenate(xyz).astype(np.float32), np.concatenate(img).astype(np.int32),
            np.concatenate(tie).astype(np.int32), np.array(co, np.int32))


def quat_angle(q1, q2):
    """rotation angle (rad) between two unit quaternions (x,y,z,w)"""
    d = abs(float(np.dot(q1 / np.linalg.norm(q1), q2 / np.linalg.norm(q2))))
    return 2.0 * np.arccos(min(1.0, d))
