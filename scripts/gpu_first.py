"""Scratch GPU bring-up: each stage against the oracle on a small case. Usage: python scripts/gpu_first.py [stage ...]"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moped_b200 import synth, capi
from oracle import oracle

stages = sys.argv[1:] or ["exact", "tensor", "cluster", "pose", "filter", "frame"]
n_obj = int(os.environ.get("NOBJ", "20"))
db = synth.make_db(n_obj, 1000)
fr = synth.make_frame(db, 2000, n_visible=4)
dbn = oracle.norm_rows(db["desc"]); qn = oracle.norm_rows(fr["desc"])
ctx = capi.Context(0)
t = time.time(); ctx.db_upload(dbn, db["xyz"], db["model_of_row"], n_obj); print("upload s", time.time() - t, flush=True)
ctx.set_cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
cams = oracle.cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
t = time.time(); oidx, odist = oracle.match_2nn(dbn, qn); print("oracle 2nn s", time.time() - t, flush=True)
om, _, _ = oracle.match(dbn, db["xyz"], db["model_of_row"], n_obj, qn, fr["xy"], fr["image_idx"], 0.8)

def cmp_match(mode, name):
    t = time.time(); r, d, a, st = ctx.match(qn, 0.8, mode); dt = time.time() - t
    print(f"[{name}] time {dt*1e3:.2f} ms  idx equal {np.array_equal(r, oidx)}  dist equal {np.array_equal(d, odist)}  stats {st}", flush=True)
    if not np.array_equal(r, oidx):
        bad = np.nonzero((r != oidx).any(1))[0]
        print("   mismatches", len(bad), bad[:10], r[bad[:5]], oidx[bad[:5]], d[bad[:5]], odist[bad[:5]])
    acc_o = odist[:, 0] / odist[:, 1] < 0.8
    print("   accepted equal", np.array_equal(a, acc_o), a.sum())

if "exact" in stages:
    cmp_match(capi.MATCH_EXACT, "exact"); cmp_match(capi.MATCH_EXACT, "exact2")
if "tensor" in stages:
    cmp_match(capi.MATCH_TENSOR, "tensor"); cmp_match(capi.MATCH_TENSOR, "tensor2")
oc = oracle.cluster(om, 1)
if "cluster" in stages:
    gc = ctx.cluster(om, 1)
    print("[cluster]", {k: np.array_equal(gc[k], oc[k]) for k in oc}, len(oc["model"]), flush=True)
# cluster points
pts_xy, pts_xyz, pts_img, co = [], [], [], [0]
for c in range(len(oc["model"])):
    mem = oc["members"][oc["offsets"][c]:oc["offsets"][c + 1]]; lo = om["offsets"][oc["model"][c]]
    pts_xy.append(om["xy"][lo + mem]); pts_xyz.append(om["xyz"][lo + mem]); pts_img.append(om["image"][lo + mem]); co.append(co[-1] + len(mem))
pts_xy = np.concatenate(pts_xy); pts_xyz = np.concatenate(pts_xyz); pts_img = np.concatenate(pts_img); co = np.array(co, np.int32)
if "pose" in stages:
    H = 64; hc, sp, iq = [], [], []
    for c in range(len(co) - 1):
        mem = oc["members"][oc["offsets"][c]:oc["offsets"][c + 1]]
        ok, pos, quat = oracle.draw_samples(pts_xy[co[c]:co[c + 1]], pts_img[co[c]:co[c + 1]], mem, 5, 1234 + c, H)
        hc += [c] * H; sp.append(pos); iq.append(quat)
    hc = np.array(hc, np.int32); sp = np.concatenate(sp); iq = np.concatenate(iq)
    P = (600, 200, 4, 5, 6, 10.0)
    t = time.time(); n_in, plm, prf, err, masks = ctx.pose_hypotheses(co, pts_xy, pts_xyz, pts_img, hc, sp, iq, P); print("[pose] gpu hyps s", time.time() - t, flush=True)
    mm = 0; acc = 0; dts = []; dqs = []; mask_bad = 0
    for h in range(len(hc)):
        c = hc[h]; s = slice(co[c], co[c + 1])
        r, olm, orf, oerr, omask = oracle.hypothesis(pts_xy[s], pts_xyz[s], pts_img[s], cams, sp[h], iq[h], 200, 10.0, 6)
        if (r > 6) != (n_in[h] > 6) or (r < 0) != (n_in[h] < 0): mm += 1
        if r > 6 and n_in[h] > 6:
            acc += 1
            dts.append(np.abs(orf[4:] - prf[h][4:]).max()); dqs.append(min(np.abs(orf[:4] - prf[h][:4]).max(), np.abs(orf[:4] + prf[h][:4]).max()))
            if not np.array_equal(omask.astype(bool), masks[h]): mask_bad += 1
    dts = np.array(dts); dqs = np.array(dqs)
    print(f"[pose] hyps {len(hc)} accept-mismatch {mm} accepted {acc} dt max {dts.max():.2e} p95 {np.percentile(dts,95):.2e} dq max {dqs.max():.2e} mask differ {mask_bad}", flush=True)
    t = time.time(); found, pose, nt = ctx.pose_ransac(co, pts_xy, pts_xyz, pts_img, P, seed=7); print("[ransac] s", time.time() - t, found, nt, flush=True)
    for task in range(len(found)):
        c = task // 4; s = slice(co[c], co[c + 1])
        mem = oc["members"][oc["offsets"][c]:oc["offsets"][c + 1]]
        seed = (7 + 0x9E3779B97F4A7C15 * (task + 1)) & 0xFFFFFFFFFFFFFFFF
        # gpu tie ids are None in the host API -> positions
        f, p, it = oracle.ransac(pts_xy[s], pts_xyz[s], pts_img[s], None, cams, P, seed)
        print("   task", task, "gpu", found[task], nt[task], "oracle", f, it, "dpose", np.abs(p - pose[task]).max())
if "filter" in stages:
    found, pose, nt = ctx.pose_ransac(co, pts_xy, pts_xyz, pts_img, (600, 200, 4, 5, 6, 10.0), seed=7)
    objm = np.repeat(oc["model"], 4)[found]; objp = pose[found]
    of = oracle.filter_objects(om, cams, objm, objp, (5, 4096.0, 2.0))
    gf = ctx.filter(om, objm, objp, (5, 4096.0, 2.0))
    print("[filter] keep eq", np.array_equal(of["keep"], gf["keep"]), "score maxdiff", np.abs(of["score"] - gf["score"]).max() if len(objm) else 0,
          "clusters eq", np.array_equal(of["offsets"], gf["offsets"]), np.array_equal(of["members"], gf["members"]), flush=True)
if "frame" in stages:
    for i in range(3):
        t = time.time(); out = ctx.process_frame(qn, fr["xy"], fr["image_idx"], want_times=True); dt = time.time() - t
        print(f"[frame] {dt*1e3:.2f} ms objects {out['model']} gt {np.sort(fr['gt_model'])} stage_ms {out['stage_ms']}", flush=True)
    for m, p in zip(out["model"], out["pose"]):
        j = list(fr["gt_model"]).index(m); g = fr["gt_pose"][j]
        print("   ", m, "dt", np.abs(p[4:] - g[4:]).max(), "dq", min(np.abs(p[:4] - g[:4]).max(), np.abs(p[:4] + g[:4]).max()))
print("launches", ctx.launches)
