"""The oracle's stages chained like Moped::processImages chains the reference's (moped2/libmoped/src/moped.cpp:166-194, stage list and
parameters of config.hpp:83-120): MATCH -> CLUSTER -> POSE -> FILTER -> POSE2 -> FILTER2, with the RANSAC tasks on the seedable streams
libmoped_cuda's frame pipeline uses (task t of a POSE call draws from seed + 0x9E3779B97F4A7C15 * (t + 1); seeds 1 and 2 for POSE and
POSE2, mc_pipeline_default_params). Test infrastructure: what mc_process_frame* must reproduce — bit for bit in exact-order mode."""
import numpy as np

from conftest import cluster_points
from oracle import oracle

GOLDEN = 0x9E3779B97F4A7C15
M64 = 0xFFFFFFFFFFFFFFFF
POSE1 = (600, 200, 4, 5, 6, 10.0)
POSE2 = (100, 500, 4, 6, 8, 5.0)
FILTER1 = (5, 4096.0, 2.0)
FILTER2 = (7, 4096.0, 3.0)


def pose_step(matches, clusters, cams, params, seed, objects):
    """POSE_RANSAC_LM_DIFF_REPROJECTION_CPU::process (:264-307): tasks = every (cluster, try), appended in task order."""
    xy, xyz, img, tie, co = cluster_points(matches, clusters)
    n_tests = []
    for task in range(len(clusters["model"]) * params[2]):
        k = task // params[2]
        s = slice(co[k], co[k + 1])
        f, p, it = oracle.ransac(xy[s], xyz[s], img[s], tie[s], cams, params, (seed + GOLDEN * (task + 1)) & M64)
        n_tests.append(it)
        if f:
            objects.append((int(clusters["model"][k]), p.copy()))
    return n_tests


def filter_step(matches, cams, objects, params):
    """FILTER_PROJECTION_CPU::process (:80-162): surviving objects in list order, their rebuilt clusters model-major."""
    if not objects:
        return [], dict(model=np.zeros(0, np.int32), offsets=np.zeros(1, np.int32), members=np.zeros(0, np.int32)), np.zeros(0, np.float32)
    om = np.array([o[0] for o in objects], np.int32)
    op = np.stack([o[1] for o in objects]).astype(np.float32)
    f = oracle.filter_objects(matches, cams, om, op, params)
    keep = f["keep"]
    surv = [objects[i] for i in range(len(objects)) if keep[i]]
    order = sorted((int(om[i]), i) for i in range(len(objects)) if keep[i])          # clusters: model-major, then list order
    cl = dict(model=np.array([m for m, _ in order], np.int32), offsets=f["offsets"], members=f["members"])
    return surv, cl, f["score"][keep]


def frame(dbn, db_xyz, model_of_row, n_models, qn, q_xy, q_image, K, cam_pose, n_images=1, seeds=(1, 2)):
    cams = oracle.cameras(K, cam_pose)
    m, idx, dist = oracle.match(dbn, db_xyz, model_of_row, n_models, qn, q_xy, q_image, 0.8)
    cl = oracle.cluster(m, n_images)
    objects = []
    pose_step(m, cl, cams, POSE1, seeds[0], objects)
    objects, cl2, _ = filter_step(m, cams, objects, FILTER1)
    pose_step(m, cl2, cams, POSE2, seeds[1], objects)
    objects, _, score = filter_step(m, cams, objects, FILTER2)
    return dict(model=np.array([o[0] for o in objects], np.int32),
                pose=np.stack([o[1] for o in objects]).astype(np.float32) if objects else np.zeros((0, 7), np.float32),
                score=score, matches=m, clusters=cl)
