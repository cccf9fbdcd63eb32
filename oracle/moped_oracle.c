/*
 * moped_oracle.c — plain-C CPU restatement of MOPED's recognition hot path
 * (match -> mean-shift cluster -> RANSAC + Levenberg-Marquardt -> projection filter).
 *
 * TEST INFRASTRUCTURE ONLY (see moped_oracle.h). Written from the reference's behaviour, not copied
 * from it; each function cites the reference lines it restates (paths relative to /root/reference/,
 * `libs.tgz!` = moped2/libmoped/libs/libs.tgz). Arithmetic is fp32 like the reference (Float=float,
 * moped2/libmoped/include/moped.hpp:74-78); built with -ffp-contract=off so that every product and
 * sum below rounds exactly where it is written.
 */
#include "moped_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>

/* ============================================================================================
 * MATCH
 * ============================================================================================ */

/* L2 normalisation of each descriptor: sequential fp32 sum of squares, 1/sqrt, scale.
 * Restates MATCH_ANN_CPU::norm, moped2/libmoped/src/match/MATCH_ANN_CPU.hpp:54-57.
 * NOTE the reference binary is built with -ffast-math, where g++ turns 1./sqrtf(x) into RSQRTSS
 * plus one Newton step (CPU-vendor dependent bits); this restatement uses the correctly rounded
 * 1/sqrtf. Parity tests therefore feed both sides descriptors normalised by one party. */
void mo_norm_rows(float *desc, int n, int D) {
	for (int i = 0; i < n; i++) {
		float *d = desc + (long)i * D;
		float s = 0.f;
		for (int x = 0; x < D; x++) s += d[x] * d[x];
		s = (float)(1. / sqrtf(s));
		for (int x = 0; x < D; x++) d[x] *= s;
	}
}

/* Exact 2 nearest neighbours under squared L2, per query.
 * Restates the reference matcher in exact mode: ANNkd_tree::annkSearch with eps=0
 * (MATCH_ANN_CPU.hpp:162 -> libs.tgz!ann_1.1.1/kd_search.cpp:89-121), equivalently
 * ANNbruteForce::annkSearch (libs.tgz!ann_1.1.1/brute.cpp:54-78). The distance is the leaf scan's
 * (kd_search.cpp:188-199): t = q[d]-p[d]; dist = dist + t*t for d = 0..D-1 in order, product and
 * sum each rounded to fp32 (the compiled reference uses VMULSS+VADDSS, no FMA), abandoning a row as
 * soon as the partial sum exceeds the current 2nd best. The 2-element queue keeps the
 * earlier-inserted row on equal keys (libs.tgz!ann_1.1.1/pr_queue_k.h:99-113); scanning rows in
 * index order, the lower row id wins an exact tie (the kd-tree visits rows in tree order, so exact
 * ties are the one case where the two may differ). */
void mo_match_2nn(const float *db, int N, int D, const float *q, int Q, int *idx2, float *dist2) {
	#pragma omp parallel for schedule(dynamic, 16)
	for (int i = 0; i < Q; i++) {
		const float *qq = q + (long)i * D;
		float k0 = FLT_MAX, k1 = FLT_MAX;          /* PQ_NULL_KEY = ANN_DIST_INF */
		int i0 = -1, i1 = -1;
		int filled = 0;
		for (int r = 0; r < N; r++) {
			const float *pp = db + (long)r * D;
			float min_dist = (filled >= 2) ? k1 : FLT_MAX;
			float dist = 0.f;
			int d;
			for (d = 0; d < D; d++) {
				float t = qq[d] - pp[d];
				float t2 = t * t;
				dist = dist + t2;
				if (dist > min_dist) break;
			}
			if (d < D) continue;
			/* insert: slide strictly larger keys up */
			if (filled >= 1 && k0 > dist) { k1 = k0; i1 = i0; k0 = dist; i0 = r; }
			else if (filled == 0) { k0 = dist; i0 = r; }
			else if (filled == 1 || k1 > dist) { k1 = dist; i1 = r; }
			if (filled < 2) filled++;
		}
		idx2[2 * i] = i0; idx2[2 * i + 1] = i1;
		dist2[2 * i] = k0; dist2[2 * i + 1] = k1;
	}
}

/* Ratio test on SQUARED distances and bucketing by model, in query order.
 * Restates MATCH_ANN_CPU::process, MATCH_ANN_CPU.hpp:165-176. Output: for each model m,
 * match_query/match_row[model_offsets[m] .. model_offsets[m+1]) = accepted queries (ascending) and
 * their nearest DB row. Returns the number of accepted matches. */
int mo_match_emit(const int *idx2, const float *dist2, int Q, float ratio, const int *model_of_row,
                  int n_models, int *match_query, int *match_row, int *model_offsets) {
	int *count = (int *)calloc(n_models + 1, sizeof(int));
	for (int i = 0; i < Q; i++)
		if (dist2[2 * i] / dist2[2 * i + 1] < ratio) count[model_of_row[idx2[2 * i]]]++;
	int t = 0;
	for (int m = 0; m < n_models; m++) { model_offsets[m] = t; t += count[m]; count[m] = model_offsets[m]; }
	model_offsets[n_models] = t;
	for (int i = 0; i < Q; i++)
		if (dist2[2 * i] / dist2[2 * i + 1] < ratio) {
			int m = model_of_row[idx2[2 * i]];
			match_query[count[m]] = i; match_row[count[m]] = idx2[2 * i]; count[m]++;
		}
	free(count);
	return t;
}

/* ============================================================================================
 * CLUSTER
 * ============================================================================================ */

/* Canopy mean-shift over n 2-D points (one model, one image).
 * Restates CLUSTER_MEAN_SHIFT_CPU::MeanShift, moped2/libmoped/src/cluster/CLUSTER_MEAN_SHIFT_CPU.hpp:80-158:
 * every point starts as a canopy; per iteration (i) each live canopy's aggregate = size-weighted mean
 * of the centres within Radius (:102-120), (ii) in list order, every earlier canopy whose aggregate
 * lies within Merge is redirected to the current one, together with its previous target (:122-132),
 * (iii) in list order, redirected canopies fold into their target (:134-148); stop when an iteration
 * merges nothing. Emits canopies with >= minpts points; members in splice order (:151-157).
 * cluster_offsets needs n+1 ints, members n ints. Returns number of clusters. */
int mo_meanshift(const float *xy, int n, float radius, float merge, int minpts, int maxiter,
                 int *cluster_offsets, int *members) {
	float sq_radius = radius * radius, sq_merge = merge * merge;
	float *cx = (float *)malloc(sizeof(float) * 4 * (n + 1));
	float *cy = cx + n, *ax = cy + n, *ay = ax + n;
	int *size = (int *)malloc(sizeof(int) * 6 * (n + 1));
	int *target = size + n, *alive = target + n, *head = alive + n, *tail = head + n, *next = tail + n;
	int n_alive = n;
	for (int i = 0; i < n; i++) {
		cx[i] = xy[2 * i]; cy[i] = xy[2 * i + 1]; size[i] = 1; target[i] = i; alive[i] = i;
		head[i] = tail[i] = i; next[i] = -1;
	}
	int done = 0;
	for (int it = 0; !done && it < maxiter; it++) {
		done = 1;
		for (int a = 0; a < n_alive; a++) {
			int c = alive[a];
			float sx = cx[c] * size[c], sy = cy[c] * size[c];
			int touch = size[c];
			for (int b = 0; b < n_alive; b++) {
				int o = alive[b];
				if (o == c) continue;
				float dx = cx[o] - cx[c], dy = cy[o] - cy[c];
				float dist = 0.f; dist += dx * dx; dist += dy * dy;
				if (dist < sq_radius) { touch += size[o]; sx += cx[o] * size[o]; sy += cy[o] * size[o]; }
			}
			ax[c] = sx / touch; ay[c] = sy / touch;
		}
		for (int a = 0; a < n_alive; a++) {
			int c = alive[a];
			for (int b = 0; b < a; b++) {
				int o = alive[b];
				float dx = ax[o] - ax[c], dy = ay[o] - ay[c];
				float dist = 0.f; dist += dx * dx; dist += dy * dy;
				if (dist < sq_merge) { target[target[o]] = c; target[o] = c; }
			}
		}
		int w = 0;
		for (int a = 0; a < n_alive; a++) {
			int c = alive[a];
			if (target[c] != c) {
				int t = target[c];
				/* list size of the target == its running size */
				cx[t] = cx[t] * size[t] + cx[c] * size[c];
				cy[t] = cy[t] * size[t] + cy[c] * size[c];
				next[tail[t]] = head[c]; tail[t] = tail[c];
				size[t] += size[c];
				cx[t] /= size[t]; cy[t] /= size[t];
				done = 0;
			} else alive[w++] = c;
		}
		n_alive = w;
	}
	int nc = 0, t = 0;
	for (int a = 0; a < n_alive; a++) {
		int c = alive[a];
		if (size[c] < minpts) continue;
		cluster_offsets[nc++] = t;
		for (int p = head[c]; p >= 0; p = next[p]) members[t++] = p;
	}
	cluster_offsets[nc] = t;
	free(cx); free(size);
	return nc;
}

/* CLUSTER step over all models: split each model's matches by image (in match order), mean-shift
 * each image's points, append clusters per model (image-major). Members are indices into the
 * model's match list. Restates CLUSTER_MEAN_SHIFT_CPU::process, CLUSTER_MEAN_SHIFT_CPU.hpp:182-199.
 * Buffers: cluster_model/cluster_offsets need (#matches+1) ints, members #matches ints. */
int mo_cluster(const int *match_offsets, const int *match_image, const float *match_xy, int n_models, int n_images,
               float radius, float merge, int minpts, int maxiter,
               int *cluster_model, int *cluster_offsets, int *members) {
	int nc = 0, t = 0;
	cluster_offsets[0] = 0;
	for (int m = 0; m < n_models; m++) {
		int lo = match_offsets[m], hi = match_offsets[m + 1], cnt = hi - lo;
		if (cnt <= 0) continue;
		float *pxy = (float *)malloc(sizeof(float) * 2 * cnt);
		int *pid = (int *)malloc(sizeof(int) * (3 * cnt + 2));
		int *coff = pid + cnt, *cmem = coff + cnt + 1;
		for (int im = 0; im < n_images; im++) {
			int k = 0;
			for (int j = lo; j < hi; j++)
				if (match_image[j] == im) { pxy[2 * k] = match_xy[2 * j]; pxy[2 * k + 1] = match_xy[2 * j + 1]; pid[k++] = j - lo; }
			if (!k) continue;
			int c = mo_meanshift(pxy, k, radius, merge, minpts, maxiter, coff, cmem);
			for (int ci = 0; ci < c; ci++) {
				cluster_model[nc] = m;
				for (int j = coff[ci]; j < coff[ci + 1]; j++) members[t++] = pid[cmem[j]];
				cluster_offsets[++nc] = t;
			}
		}
		free(pxy); free(pid);
	}
	return nc;
}

/* ============================================================================================
 * POSE
 * ============================================================================================ */

/* Rotation (unit quaternion x,y,z,w) + translation -> 3x4, row-major.
 * Restates TransformMatrix::init, moped2/libmoped/include/moped.hpp:175-182. */
static void tm_init(float *T, const float *q, const float *t) {
	T[0] = 1 - 2 * q[1] * q[1] - 2 * q[2] * q[2]; T[1] = 2 * q[0] * q[1] - 2 * q[3] * q[2]; T[2] = 2 * q[0] * q[2] + 2 * q[3] * q[1]; T[3] = t[0];
	T[4] = 2 * q[0] * q[1] + 2 * q[3] * q[2]; T[5] = 1 - 2 * q[0] * q[0] - 2 * q[2] * q[2]; T[6] = 2 * q[1] * q[2] - 2 * q[3] * q[0]; T[7] = t[1];
	T[8] = 2 * q[0] * q[2] - 2 * q[3] * q[1]; T[9] = 2 * q[1] * q[2] + 2 * q[3] * q[0]; T[10] = 1 - 2 * q[0] * q[0] - 2 * q[1] * q[1]; T[11] = t[2];
}
/* moped.hpp:183-188 */
static void tm_transform(const float *T, float *d, const float *o) {
	float x = o[0] * T[0] + o[1] * T[1] + o[2] * T[2] + T[3];
	float y = o[0] * T[4] + o[1] * T[5] + o[2] * T[6] + T[7];
	float z = o[0] * T[8] + o[1] * T[9] + o[2] * T[10] + T[11];
	d[0] = x; d[1] = y; d[2] = z;
}
/* moped.hpp:190-200 */
static void tm_inverse(const float *T, float *d, const float *o) {
	float a = o[0] - T[3], b = o[1] - T[7], c = o[2] - T[11];
	d[0] = a * T[0] + b * T[4] + c * T[8];
	d[1] = a * T[1] + b * T[5] + c * T[9];
	d[2] = a * T[2] + b * T[6] + c * T[10];
}
/* Pt<4>::norm, moped.hpp:122: fp32 sum of squares, fp32 sqrt, reciprocal narrowed to fp32. */
static void quat_norm(float *q) {
	float d = 0.f;
	for (int x = 0; x < 4; x++) d += q[x] * q[x];
	d = (float)(1. / (double)sqrtf(d));
	for (int x = 0; x < 4; x++) q[x] *= d;
}

void mo_camera_init(mo_camera *cam, const float *K4, const float *cam_pose7) {
	memcpy(cam->K, K4, 4 * sizeof(float));
	tm_init(cam->TM, cam_pose7, cam_pose7 + 4);   /* image->TM.init(cameraPose), moped.cpp:168-169 */
}

/* The seedable stand-in for libc rand() used by oracle/ref_harness.cpp (same LCG). */
int mo_rand(uint64_t *state) {
	*state = *state * 6364136223846793005ULL + 1442695040888963407ULL;
	return (int)((*state >> 33) & 0x7fffffffULL);
}

typedef struct { float key; int tie; int pos; } mo_keyed;
static int keyed_cmp(const void *a, const void *b) {
	const mo_keyed *x = (const mo_keyed *)a, *y = (const mo_keyed *)b;
	if (x->key != y->key) return x->key < y->key ? -1 : 1;
	return (x->tie > y->tie) - (x->tie < y->tie);
}

/* Random sample of n_samples distinct (image, coord2D) points: key every cluster point with
 * (float)rand() in cluster order, sort ascending (ties -> LmData address = position `tie`), walk
 * from the front skipping points whose (image, coord2D) was already taken.
 * Restates POSE_RANSAC_LM_DIFF_REPROJECTION_CPU::randSample, .../pose/POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp:76-98.
 * tie_ids[i] = the point's index in its model's match list (the reference breaks key ties by
 * LmData address, i.e. match index); NULL = cluster position. Returns 1 on success. */
int mo_rand_sample(uint64_t *state, const float *xy, const int *image, const int *tie_ids, int n, int n_samples, int *sample_pos) {
	mo_keyed *k = (mo_keyed *)malloc(sizeof(mo_keyed) * (n > 0 ? n : 1));
	for (int i = 0; i < n; i++) { k[i].key = (float)mo_rand(state); k[i].tie = tie_ids ? tie_ids[i] : i; k[i].pos = i; }
	qsort(k, n, sizeof(mo_keyed), keyed_cmp);
	int used = 0, ns = 0;
	for (int f = 0; f < n && used < n_samples; f++) {
		int p = k[f].pos, dup = 0;
		for (int j = 0; j < ns; j++) {
			int s = sample_pos[j];
			if (image[s] == image[p] && xy[2 * s] == xy[2 * p] && xy[2 * s + 1] == xy[2 * p + 1]) { dup = 1; break; }
		}
		if (!dup) { sample_pos[ns++] = p; used++; }
	}
	free(k);
	return used == n_samples;
}

/* Initial pose: quaternion components (rand()&255)/256, translation (0,0,0.5).
 * Restates initPose, POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp:182-186. g++ evaluates the four
 * rand() arguments right to left on x86-64, so w draws first (pinned by tests against the build). */
void mo_init_pose(uint64_t *state, float *pose7) {
	for (int j = 3; j >= 0; j--) pose7[j] = (float)((mo_rand(state) & 255) / 256.);
	pose7[4] = 0.f; pose7[5] = 0.f; pose7[6] = 0.5f;
}

/* Residual vector of a pose (7 = raw quaternion + t) over n_pts 2D-3D correspondences:
 * normalise q, R*X+t, world->camera, pinhole; residuals are the SQUARED pixel differences
 * ((u-u0)^2, (v-v0)^2), or (-z+10, -z+10) for a point behind the camera.
 * Restates lmFuncQuat, POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp:100-138. */
void mo_lm_func(const float *p7, float *res, int n_pts, const float *xy, const float *xyz, const int *image, const mo_camera *cams) {
	float q[4] = { p7[0], p7[1], p7[2], p7[3] };
	quat_norm(q);
	float T[12];
	tm_init(T, q, p7 + 4);
	for (int i = 0; i < n_pts; i++) {
		const mo_camera *cam = &cams[image[i]];
		float p3[3];
		tm_transform(T, p3, xyz + 3 * i);
		tm_inverse(cam->TM, p3, p3);
		float u = p3[0] / p3[2] * cam->K[0] + cam->K[2];
		float v = p3[1] / p3[2] * cam->K[1] + cam->K[3];
		if (p3[2] < 0) {
			res[2 * i] = -p3[2] + 10;
			res[2 * i + 1] = -p3[2] + 10;
		} else {
			float a = u - xy[2 * i], b = v - xy[2 * i + 1];
			res[2 * i] = a * a;
			res[2 * i + 1] = b * b;
		}
	}
}

/* e = -y, returns ||e||^2 with levmar's 4-accumulator, 8-way unrolled, downward order.
 * Restates slevmar_L2nrmxmy (x == zero vector), libs.tgz!levmar-2.4/misc_core.c:712-790. */
static float l2_neg(float *e, const float *y, int n) {
	float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
	int blockn = (n >> 3) << 3;
	for (int i = blockn - 1; i > 0; i -= 8) {
		e[i] = 0.f - y[i]; s0 += e[i] * e[i];
		e[i - 1] = 0.f - y[i - 1]; s1 += e[i - 1] * e[i - 1];
		e[i - 2] = 0.f - y[i - 2]; s2 += e[i - 2] * e[i - 2];
		e[i - 3] = 0.f - y[i - 3]; s3 += e[i - 3] * e[i - 3];
		e[i - 4] = 0.f - y[i - 4]; s0 += e[i - 4] * e[i - 4];
		e[i - 5] = 0.f - y[i - 5]; s1 += e[i - 5] * e[i - 5];
		e[i - 6] = 0.f - y[i - 6]; s2 += e[i - 6] * e[i - 6];
		e[i - 7] = 0.f - y[i - 7]; s3 += e[i - 7] * e[i - 7];
	}
	for (int i = blockn; i < n; i++) { e[i] = 0.f - y[i]; s0 += e[i] * e[i]; }
	return s0 + s1 + s2 + s3;
}

/* Solve A x = B (m x m) by Crout LU with implicit row scaling and partial pivoting.
 * Restates sAx_eq_b_LU_noLapack, libs.tgz!levmar-2.4/Axb_core.c:888-1035. Returns 0 if singular. */
#define LM_M 7
static int lu_solve(const float *A, const float *B, float *x, int m) {
	float a[LM_M * LM_M], work[LM_M];
	int idx[LM_M], maxi = -1;
	for (int i = 0; i < m * m; i++) a[i] = A[i];
	for (int i = 0; i < m; i++) x[i] = B[i];
	for (int i = 0; i < m; i++) {
		float mx = 0.f;
		for (int j = 0; j < m; j++) { float t = fabsf(a[i * m + j]); if (t > mx) mx = t; }
		if (mx == 0.f) return 0;
		work[i] = 1.0f / mx;
	}
	for (int j = 0; j < m; j++) {
		for (int i = 0; i < j; i++) {
			float sum = a[i * m + j];
			for (int k = 0; k < i; k++) sum -= a[i * m + k] * a[k * m + j];
			a[i * m + j] = sum;
		}
		float mx = 0.f;
		for (int i = j; i < m; i++) {
			float sum = a[i * m + j];
			for (int k = 0; k < j; k++) sum -= a[i * m + k] * a[k * m + j];
			a[i * m + j] = sum;
			float t = work[i] * fabsf(sum);
			if (t >= mx) { mx = t; maxi = i; }
		}
		/* levmar initialises maxi = -1 and a first column of NaNs (0 * inf after an overflowed J^T J) never sets it: the reference
		 * then swaps with the row BEFORE its matrix — undefined behaviour, whatever happens to lie on its stack decides. No
		 * reference result exists for that input; the restatement (and the CUDA LM) call the system singular instead. */
		if (maxi < 0) return 0;
		if (j != maxi) {
			for (int k = 0; k < m; k++) { float t = a[maxi * m + k]; a[maxi * m + k] = a[j * m + k]; a[j * m + k] = t; }
			work[maxi] = work[j];
		}
		idx[j] = maxi;
		if (a[j * m + j] == 0.f) a[j * m + j] = FLT_EPSILON;
		if (j != m - 1) {
			float t = 1.0f / a[j * m + j];
			for (int i = j + 1; i < m; i++) a[i * m + j] *= t;
		}
	}
	int k = 0;
	for (int i = 0; i < m; i++) {
		int j = idx[i];
		float sum = x[j];
		x[j] = x[i];
		if (k != 0) { for (j = k - 1; j < i; j++) sum -= a[i * m + j] * x[j]; }
		else if (sum != 0.f) k = i + 1;
		x[i] = sum;
	}
	for (int i = m - 1; i >= 0; i--) {
		float sum = x[i];
		for (int j = i + 1; j < m; j++) sum -= a[i * m + j] * x[j];
		x[i] = sum / a[i * m + i];
	}
	return 1;
}

/* Levenberg-Marquardt with forward-difference Jacobian and Broyden rank-1 updates, target vector 0,
 * default options (tau=1e-3, eps1=eps2=eps3=1e-17, delta=1e-6), m=7 parameters, n=2*n_pts residuals.
 * Restates slevmar_dif as called by optimizeCamera (POSE_..._CPU.hpp:153-154):
 * libs.tgz!levmar-2.4/lm_core.c:427-836, Jacobian misc_core.c:135-168, defaults lm.h:83-85.
 * The reference build folds LM_FINITE() to true (-ffinite-math-only), so there is no stop=7.
 * Returns the iteration count, or -1 (LM_ERROR) on stop=4. info10 as levmar's info[]. */
typedef void (*lm_fn)(const float *p7, float *res, const void *ctx);      /* residual callback: levmar's `func` + adata */

/* levmar's `if(!LM_FINITE(...)) stop=7` (lm_core.c:551,732). The reference is built with -ffast-math (-ffinite-math-only), which
 * folds the test away — that is the default here (0). A strict-IEEE build of the same sources keeps it: tests that compare with
 * such a build bit for bit switch it on (tests/test_oracle3d_pose.py). */
static int g_lm_finite_check = 0;
void mo_set_lm_finite_check(int on) { g_lm_finite_check = on != 0; }

static int levmar_dif_fn(float *p, int n, int itmax, lm_fn fn, const void *ctx, float *info) {
	const int m = LM_M;
	const float tau = 1E-03f, eps1 = 1E-17f, eps2 = 1E-17f, eps2_sq = 1E-17f * 1E-17f, eps3 = 1E-17f, delta = 1E-06f;
	float *buf = (float *)malloc(sizeof(float) * (size_t)(5 * n + n * m));
	float *e = buf, *hx = e + n, *wrk = hx + n, *wrk2 = wrk + n, *hxx = wrk2 + n, *jac = hxx + n;
	float jtj[LM_M * LM_M], jte[LM_M], Dp[LM_M], diag[LM_M], pDp[LM_M];
	float mu = 0.f, jte_inf = 0.f, p_L2 = 0.f, Dp_L2 = FLT_MAX, p_eL2, pDp_eL2, init_eL2, tmp;
	int nu = 20, nu2, stop = 0, nfev, njap = 0, nlss = 0, K = 10, updjac = 0, updp = 1, newjac = 0, k;
	/* the reference executable runs with FTZ|DAZ (crtfastmath): the dominant
	 * exit, stop=2, is an underflow event, so flush denormals here too */
	unsigned csr_saved = _mm_getcsr();
	_mm_setcsr(csr_saved | 0x8040u);

	fn(p, hx, ctx); nfev = 1;
	p_eL2 = l2_neg(e, hx, n);
	init_eL2 = p_eL2;
	if (g_lm_finite_check && !isfinite(p_eL2)) stop = 7;

	for (k = 0; k < itmax && !stop; ++k) {
		if (p_eL2 <= eps3) { stop = 6; break; }

		if ((updp && nu > 16) || updjac == K) {
			for (int j = 0; j < m; j++) {
				float d = 1E-04f * p[j];
				d = fabsf(d);
				if (d < delta) d = delta;
				float save = p[j];
				p[j] += d;
				fn(p, hxx, ctx);
				p[j] = save;
				d = 1.0f / d;
				for (int i = 0; i < n; i++) jac[i * m + j] = (hxx[i] - hx[i]) * d;
			}
			++njap; nfev += m;
			nu = 2; updjac = 0; updp = 0; newjac = 1;
		}

		if (newjac) {
			newjac = 0;
			for (int i = 0; i < m * m; i++) jtj[i] = 0.f;
			for (int i = 0; i < m; i++) jte[i] = 0.f;
			for (int l = n - 1; l >= 0; l--) {
				const float *jl = jac + l * m;
				for (int i = m - 1; i >= 0; i--) {
					float alpha = jl[i];
					for (int j = i; j >= 0; j--) jtj[i * m + j] += jl[j] * alpha;
					jte[i] += alpha * e[l];
				}
			}
			for (int i = m - 1; i >= 0; i--)
				for (int j = i + 1; j < m; j++) jtj[i * m + j] = jtj[j * m + i];
			p_L2 = jte_inf = 0.f;
			for (int i = 0; i < m; i++) {
				tmp = fabsf(jte[i]);
				if (jte_inf < tmp) jte_inf = tmp;
				diag[i] = jtj[i * m + i];
				p_L2 += p[i] * p[i];
			}
		}

		if (jte_inf <= eps1) { Dp_L2 = 0.f; stop = 1; break; }

		if (k == 0) {
			tmp = -FLT_MAX;
			for (int i = 0; i < m; i++) if (diag[i] > tmp) tmp = diag[i];
			mu = tau * tmp;
		}

		for (int i = 0; i < m; i++) jtj[i * m + i] += mu;

		int solved = lu_solve(jtj, jte, Dp, m); ++nlss;
		if (solved) {
			Dp_L2 = 0.f;
			for (int i = 0; i < m; i++) { tmp = Dp[i]; pDp[i] = p[i] + tmp; Dp_L2 += tmp * tmp; }
			if (Dp_L2 <= eps2_sq * p_L2) { stop = 2; break; }
			if (Dp_L2 >= (p_L2 + eps2) / (1E-12f * 1E-12f)) { stop = 4; break; }

			fn(pDp, wrk, ctx); ++nfev;
			pDp_eL2 = l2_neg(wrk2, wrk, n);
			if (g_lm_finite_check && !isfinite(pDp_eL2)) { stop = 7; break; }
			float dF = p_eL2 - pDp_eL2;
			if (updp || dF > 0) {
				for (int i = 0; i < n; i++) {
					tmp = 0.f;
					for (int l = 0; l < m; l++) tmp += jac[i * m + l] * Dp[l];
					tmp = (wrk[i] - hx[i] - tmp) / Dp_L2;
					for (int j = 0; j < m; j++) jac[i * m + j] += tmp * Dp[j];
				}
				++updjac; newjac = 1;
			}
			float dL = 0.f;
			for (int i = 0; i < m; i++) dL += Dp[i] * (mu * Dp[i] + jte[i]);
			if (dL > 0.f && dF > 0.f) {
				tmp = 2.0f * dF / dL - 1.0f;
				tmp = 1.0f - tmp * tmp * tmp;
				mu = mu * ((tmp >= 0.3333333334f) ? tmp : 0.3333333334f);
				nu = 2;
				for (int i = 0; i < m; i++) p[i] = pDp[i];
				for (int i = 0; i < n; i++) { e[i] = wrk2[i]; hx[i] = wrk[i]; }
				p_eL2 = pDp_eL2;
				updp = 1;
				continue;
			}
		}
		mu *= nu;
		nu2 = nu << 1;
		if (nu2 <= nu) { stop = 5; break; }
		nu = nu2;
		for (int i = 0; i < m; i++) jtj[i * m + i] = diag[i];
	}
	if (k >= itmax) stop = 3;
	if (info) {
		info[0] = init_eL2; info[1] = p_eL2; info[2] = jte_inf; info[3] = Dp_L2; info[4] = 0.f;
		info[5] = (float)k; info[6] = (float)stop; info[7] = (float)nfev; info[8] = (float)njap; info[9] = (float)nlss;
	}
	free(buf);
	_mm_setcsr(csr_saved);
	return (stop != 4 && stop != 7) ? k : -1;
}

typedef struct { int n_pts; const float *xy, *xyz; const int *image; const mo_camera *cams; } reproj_ctx;
static void reproj_fn(const float *p7, float *res, const void *c) {
	const reproj_ctx *r = (const reproj_ctx *)c;
	mo_lm_func(p7, res, r->n_pts, r->xy, r->xyz, r->image, r->cams);
}

int mo_levmar_dif(float *p, int n_pts, int itmax, const float *xy, const float *xyz, const int *image,
                  const mo_camera *cams, float *info) {
	reproj_ctx c = { n_pts, xy, xyz, image, cams };
	return levmar_dif_fn(p, 2 * n_pts, itmax, reproj_fn, &c, info);
}

/* LM from `pose7` over the given correspondences; on success overwrites pose7 with the solution
 * (quaternion re-normalised) and returns ||e||^2; on LM_ERROR returns -1 and leaves pose7 untouched.
 * Restates optimizeCamera, POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp:140-164. */
float mo_optimize_camera(float *pose7, int n_pts, int itmax, const float *xy, const float *xyz, const int *image, const mo_camera *cams) {
	float p[7], info[10];
	memcpy(p, pose7, sizeof p);
	int r = mo_levmar_dif(p, n_pts, itmax, xy, xyz, image, cams, info);
	if (r < 0) return (float)r;
	memcpy(pose7, p, sizeof p);
	quat_norm(pose7);
	return info[1];
}

/* Pinhole projection of a model point under `pose7` into a camera; (FLT_MAX, FLT_MAX) if z < 0.001.
 * Restates project(), moped2/libmoped/include/moped.hpp:330-354. */
void mo_project(const float *pose7, const float *xyz3, const mo_camera *cam, float *uv2) {
	float T[12], p3[3];
	tm_init(T, pose7, pose7 + 4);
	tm_transform(T, p3, xyz3);
	tm_inverse(cam->TM, p3, p3);
	uv2[0] = FLT_MAX; uv2[1] = FLT_MAX;
	if (p3[2] < 0.001) return;
	uv2[0] = p3[0] / p3[2] * cam->K[0] + cam->K[2];
	uv2[1] = p3[1] / p3[2] * cam->K[1] + cam->K[3];
}

/* Inlier test of every cluster point: squared reprojection error < err_thr (px^2).
 * Restates testAllPoints, POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp:166-180. Returns #inliers. */
int mo_test_all_points(const float *pose7, int n, const float *xy, const float *xyz, const int *image,
                       const mo_camera *cams, float err_thr, unsigned char *mask) {
	int c = 0;
	for (int i = 0; i < n; i++) {
		float uv[2];
		mo_project(pose7, xyz + 3 * i, &cams[image[i]], uv);
		float a = uv[0] - xy[2 * i], b = uv[1] - xy[2 * i + 1];
		float err = a * a + b * b;
		mask[i] = err < err_thr;
		c += mask[i];
	}
	return c;
}

static int gather(const unsigned char *mask, const int *pos, int cnt, const float *xy, const float *xyz, const int *image,
                  float *gxy, float *gxyz, int *gim, int n) {
	int k = 0;
	if (pos) { for (int j = 0; j < cnt; j++) { int i = pos[j]; gxy[2 * k] = xy[2 * i]; gxy[2 * k + 1] = xy[2 * i + 1]; memcpy(gxyz + 3 * k, xyz + 3 * i, 12); gim[k] = image[i]; k++; } }
	else { for (int i = 0; i < n; i++) if (mask[i]) { gxy[2 * k] = xy[2 * i]; gxy[2 * k + 1] = xy[2 * i + 1]; memcpy(gxyz + 3 * k, xyz + 3 * i, 12); gim[k] = image[i]; k++; } }
	return k;
}

/* One RANSAC iteration body on an explicit (sample set, initial quaternion): LM on the samples,
 * inlier test on the whole cluster, LM refit on the inliers when #inliers > min_npts.
 * Restates the loop body of RANSAC(), POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp:191-208.
 * Returns -1 if the sample fit failed (iteration skipped), else #inliers. */
int mo_hypothesis(int n, const float *xy, const float *xyz, const int *image, const mo_camera *cams,
                  const int *sample_pos, int n_samples, const float *init_quat, int max_lm, float err_thr, int min_npts,
                  float *pose_lm, float *pose_refit, float *lm_err2, unsigned char *mask) {
	float *gxy = (float *)malloc(sizeof(float) * 5 * (n + n_samples));
	float *gxyz = gxy + 2 * (n + n_samples);
	int *gim = (int *)malloc(sizeof(int) * (n + n_samples));
	float pose[7] = { init_quat[0], init_quat[1], init_quat[2], init_quat[3], 0.f, 0.f, 0.5f };
	int k = gather(NULL, sample_pos, n_samples, xy, xyz, image, gxy, gxyz, gim, n);
	float r = mo_optimize_camera(pose, k, max_lm, gxy, gxyz, gim, cams);
	lm_err2[0] = r; lm_err2[1] = -2.f;
	memset(mask, 0, n);
	int ret = -1;
	if ((int)r != -1) {
		memcpy(pose_lm, pose, sizeof pose);
		ret = mo_test_all_points(pose, n, xy, xyz, image, cams, err_thr, mask);
		if (ret > min_npts) {
			k = gather(mask, NULL, 0, xy, xyz, image, gxy, gxyz, gim, n);
			lm_err2[1] = mo_optimize_camera(pose, k, max_lm, gxy, gxyz, gim, cams);
		}
		memcpy(pose_refit, pose, sizeof pose);
	}
	free(gxy); free(gim);
	return ret;
}

/* Sequential RANSAC with early exit on the first hypothesis whose inlier count exceeds min_npts.
 * Restates RANSAC(), POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp:188-211. Returns found (0/1);
 * *iters = number of iterations consumed. */
int mo_ransac(uint64_t *state, int n, const float *xy, const float *xyz, const int *image, const int *tie_ids, const mo_camera *cams,
              int max_ransac, int max_lm, int n_pts_align, int min_npts, float err_thr, float *pose7, int *iters) {
	int *pos = (int *)malloc(sizeof(int) * (n_pts_align + 1));
	unsigned char *mask = (unsigned char *)malloc(n + 1);
	float pose_lm[7], pose_refit[7], err2[2], init[7];
	int found = 0, it;
	for (it = 0; it < max_ransac; it++) {
		if (!mo_rand_sample(state, xy, image, tie_ids, n, n_pts_align, pos)) break;
		mo_init_pose(state, init);
		int r = mo_hypothesis(n, xy, xyz, image, cams, pos, n_pts_align, init, max_lm, err_thr, min_npts, pose_lm, pose_refit, err2, mask);
		if (r < 0) { memcpy(pose7, init, sizeof init); continue; }
		memcpy(pose7, pose_refit, sizeof pose_refit);
		if (r > min_npts) { found = 1; it++; break; }
	}
	if (iters) *iters = it;
	free(pos); free(mask);
	return found;
}

/* ============================================================================================
 * FILTER
 * ============================================================================================ */

typedef struct { int image; float x, y; int match; } mo_key;
static int key_cmp(const void *a, const void *b) {
	const mo_key *p = (const mo_key *)a, *q = (const mo_key *)b;
	if (p->image != q->image) return p->image < q->image ? -1 : 1;
	if (p->x != q->x) return p->x < q->x ? -1 : 1;
	if (p->y != q->y) return p->y < q->y ? -1 : 1;
	return 0;
}

/* Projection filter: score every object by reprojecting all matches of its model (those within
 * feat_dist px^2 form its cluster, score = sum 1/(err+1)); every distinct (coord2D, image) is owned
 * by the object with the strictly highest score seen (first wins ties, objects visited model-major in
 * list order); clusters are rebuilt from owned matches of the object's own model; objects with
 * fewer than min_points owned matches or score < min_score are dropped.
 * Restates FILTER_PROJECTION_CPU::process, moped2/libmoped/src/filter/FILTER_PROJECTION_CPU.hpp:80-162.
 * Outputs: keep[n_obj], score[n_obj]; clusters (CSR over the SURVIVORS ordered model-major then
 * list order, members = indices into the model's match list). Returns #survivors. */
int mo_filter(int n_models, const int *match_offsets, const int *match_image, const float *match_xy, const float *match_xyz,
              const mo_camera *cams, int n_obj, const int *obj_model, const float *obj_pose, int min_points, float feat_dist,
              float min_score, unsigned char *keep, float *score, int *cluster_offsets, int *members) {
	int M = match_offsets[n_models];
	/* distinct (image, coord2D) keys -> key id per match */
	mo_key *keys = (mo_key *)malloc(sizeof(mo_key) * (M + 1));
	int *key_of = (int *)malloc(sizeof(int) * (M + 1));
	for (int j = 0; j < M; j++) { keys[j].image = match_image[j]; keys[j].x = match_xy[2 * j]; keys[j].y = match_xy[2 * j + 1]; keys[j].match = j; }
	qsort(keys, M, sizeof(mo_key), key_cmp);
	int nk = 0;
	for (int j = 0; j < M; j++) {
		if (j == 0 || key_cmp(&keys[j - 1], &keys[j]) != 0) nk++;
		key_of[keys[j].match] = nk - 1;
	}
	float *best_score = (float *)calloc(nk + 1, sizeof(float));
	int *best_obj = (int *)malloc(sizeof(int) * (nk + 1));
	for (int i = 0; i < nk; i++) best_obj[i] = -1;
	unsigned char *in_cl = (unsigned char *)malloc(M + 1);

	for (int m = 0; m < n_models; m++)
		for (int o = 0; o < n_obj; o++) {
			if (obj_model[o] != m) continue;
			float s = 0.f;
			for (int j = match_offsets[m]; j < match_offsets[m + 1]; j++) {
				float uv[2];
				mo_project(obj_pose + 7 * o, match_xyz + 3 * j, &cams[match_image[j]], uv);
				float a = uv[0] - match_xy[2 * j], b = uv[1] - match_xy[2 * j + 1];
				float err = a * a + b * b;
				in_cl[j] = err < feat_dist;
				if (in_cl[j]) s += 1. / (err + 1.);
			}
			score[o] = s;
			for (int j = match_offsets[m]; j < match_offsets[m + 1]; j++)
				if (in_cl[j] && best_score[key_of[j]] < s) { best_score[key_of[j]] = s; best_obj[key_of[j]] = o; }
		}

	int *owned = (int *)calloc(n_obj + 1, sizeof(int));
	for (int m = 0; m < n_models; m++)
		for (int j = match_offsets[m]; j < match_offsets[m + 1]; j++) {
			int o = best_obj[key_of[j]];
			if (o >= 0 && obj_model[o] == m) owned[o]++;
		}
	int ns = 0, t = 0;
	cluster_offsets[0] = 0;
	for (int o = 0; o < n_obj; o++) keep[o] = 0;
	for (int m = 0; m < n_models; m++)
		for (int o = 0; o < n_obj; o++) {
			if (obj_model[o] != m) continue;
			if (owned[o] < min_points || score[o] < min_score) continue;
			keep[o] = 1;
			for (int j = match_offsets[m]; j < match_offsets[m + 1]; j++)
				if (best_obj[key_of[j]] == o) members[t++] = j - match_offsets[m];
			cluster_offsets[++ns] = t;
		}
	free(keys); free(key_of); free(best_score); free(best_obj); free(in_cl); free(owned);
	return ns;
}

/* ==== moped3d: FILTER_PROJECTION_DEPTH_CPU (moped3d/libmoped/src/filter/FILTER_PROJECTION_DEPTH_CPU.hpp:50-331) =================
 * The projection filter above plus a penalty from the depth map ("incorrect score", :209-269): the model's TEST POINTS (all its
 * keypoints, or TestSampleSize of them drawn once by randSample, :77-118) are transformed by the object's pose into the depth camera;
 * a point that lands on a pixel with measured depth (fill distance <= 0) and is not occluded (measured depth >= its own) adds
 * 1 - 1/(1 + ((z - measured) / (DepthFraction * measured))^2). If more than MinKeypointFraction of the test points were usable, the sum
 * is rescaled by (#matches within PlausibleSqDistance) / (#usable points) and subtracted from the projection score. Ownership of the
 * image features is decided by the projection score WITHOUT the penalty (:276-289, `point.first < score`), pruning by the score
 * with it (:314). Pinned against the class itself compiled into oracle/_ref/libmoped3d_ref_strict.so (ref3d_harness.cpp,
 * tests/test_oracle3d_filter.py). Compiled without -fsingle-precision-constant: the literals are double. */

/* randSample (:77-92): keys (Float)rand() in keypoint order, pairs sorted by (key, index), the first n_samples indices in that order */
typedef struct { float key; int idx; } mo_skey;
static int skey_cmp(const void *a, const void *b) {
	const mo_skey *p = (const mo_skey *)a, *q = (const mo_skey *)b;
	if (p->key != q->key) return p->key < q->key ? -1 : 1;
	return p->idx < q->idx ? -1 : (p->idx > q->idx ? 1 : 0);
}
int mo_filter_depth_select(uint64_t *state, int n_keypoints, int sample_size, int *out_idx) {
	if (n_keypoints <= sample_size) { for (int i = 0; i < n_keypoints; i++) out_idx[i] = i; return n_keypoints; }
	mo_skey *k = (mo_skey *)malloc(sizeof(mo_skey) * (size_t)n_keypoints);
	for (int i = 0; i < n_keypoints; i++) { k[i].key = (float)mo_rand(state); k[i].idx = i; }
	qsort(k, n_keypoints, sizeof(mo_skey), skey_cmp);
	for (int i = 0; i < sample_size; i++) out_idx[i] = k[i].idx;
	free(k);
	return sample_size;
}

/* the penalty of one object: test points [0, n_test) of its model; depth / fill_distance are width x height row-major planes
 * (Image::getDepth / getProb). *used = usable test points. */
float mo_filter_depth_penalty(const float *pose7, int n_test, const float *test_xyz, const mo_camera *depth_cam, int width, int height,
                              const float *depth, const float *fill_distance, float depth_fraction, int *used) {
	float T[12], IS = 0.0;
	int n_used = 0;
	tm_init(T, pose7, pose7 + 4);
	for (int k = 0; k < n_test; k++) {
		float p3[3], p2[2];
		tm_transform(T, p3, test_xyz + 3 * k);
		tm_inverse(depth_cam->TM, p3, p3);
		p2[0] = p3[0] / p3[2] * depth_cam->K[0] + depth_cam->K[2];
		p2[1] = p3[1] / p3[2] * depth_cam->K[1] + depth_cam->K[3];
		/* (int) of a NaN, an infinity or a value beyond int range is the x86 "integer indefinite" 0x80000000: negative, off the image */
		if (!(p2[0] > -2147483648.f && p2[0] < 2147483648.f) || !(p2[1] > -2147483648.f && p2[1] < 2147483648.f)) continue;
		int ix = (int)p2[0], iy = (int)p2[1];
		if (ix < 0 || ix >= width || iy < 0 || iy >= height) continue;
		float distance = fill_distance[(size_t)iy * width + ix];
		if (distance > 0) continue;
		n_used++;
		float kinect = depth[(size_t)iy * width + ix], putative = p3[2];
		if (kinect < putative) continue;
		float scale = depth_fraction * kinect;
		float term = (putative - kinect) / scale;
		term *= term;
		IS += 1.0 - (1.0 / (1.0 + term));
	}
	*used = n_used;
	return IS;
}

int mo_filter_depth(int n_models, const int *match_offsets, const int *match_image, const float *match_xy, const float *match_xyz,
                    const mo_camera *cams, int n_obj, const int *obj_model, const float *obj_pose, int min_points, float feat_dist,
                    float plausible_dist, float min_score, float depth_fraction, float min_keypoint_fraction,
                    const int *test_offsets, const float *test_xyz, const mo_camera *depth_cam, int width, int height,
                    const float *depth, const float *fill_distance,
                    unsigned char *keep, float *score, int *cluster_offsets, int *members) {
	int M = match_offsets[n_models];
	mo_key *keys = (mo_key *)malloc(sizeof(mo_key) * (M + 1));
	int *key_of = (int *)malloc(sizeof(int) * (M + 1));
	for (int j = 0; j < M; j++) { keys[j].image = match_image[j]; keys[j].x = match_xy[2 * j]; keys[j].y = match_xy[2 * j + 1]; keys[j].match = j; }
	qsort(keys, M, sizeof(mo_key), key_cmp);
	int nk = 0;
	for (int j = 0; j < M; j++) {
		if (j == 0 || key_cmp(&keys[j - 1], &keys[j]) != 0) nk++;
		key_of[keys[j].match] = nk - 1;
	}
	float *best_score = (float *)calloc(nk + 1, sizeof(float));
	int *best_obj = (int *)malloc(sizeof(int) * (nk + 1));
	for (int i = 0; i < nk; i++) best_obj[i] = -1;
	unsigned char *in_cl = (unsigned char *)malloc(M + 1);

	for (int m = 0; m < n_models; m++)
		for (int o = 0; o < n_obj; o++) {
			if (obj_model[o] != m) continue;
			float s = 0.f;
			int cluster_size = 0;
			for (int j = match_offsets[m]; j < match_offsets[m + 1]; j++) {
				float uv[2];
				mo_project(obj_pose + 7 * o, match_xyz + 3 * j, &cams[match_image[j]], uv);
				float a = uv[0] - match_xy[2 * j], b = uv[1] - match_xy[2 * j + 1];
				float err = a * a + b * b;
				in_cl[j] = err < feat_dist;
				if (in_cl[j]) s += 1. / (err + 1.);
				if (err < plausible_dist) cluster_size++;
			}
			const int n_test = test_offsets[m + 1] - test_offsets[m];
			int used = 0;
			float IS = mo_filter_depth_penalty(obj_pose + 7 * o, n_test, test_xyz + 3 * (size_t)test_offsets[m], depth_cam, width, height, depth,
			                                   fill_distance, depth_fraction, &used);
			if (used <= (int)(min_keypoint_fraction * n_test)) IS = 0;
			else IS *= ((float)cluster_size) / used;
			score[o] = s - IS;
			for (int j = match_offsets[m]; j < match_offsets[m + 1]; j++)
				if (in_cl[j] && best_score[key_of[j]] < s) { best_score[key_of[j]] = s; best_obj[key_of[j]] = o; }
		}

	int *owned = (int *)calloc(n_obj + 1, sizeof(int));
	for (int m = 0; m < n_models; m++)
		for (int j = match_offsets[m]; j < match_offsets[m + 1]; j++) {
			int o = best_obj[key_of[j]];
			if (o >= 0 && obj_model[o] == m) owned[o]++;
		}
	int ns = 0, t = 0;
	cluster_offsets[0] = 0;
	for (int o = 0; o < n_obj; o++) keep[o] = 0;
	for (int m = 0; m < n_models; m++)
		for (int o = 0; o < n_obj; o++) {
			if (obj_model[o] != m) continue;
			if (owned[o] < min_points || score[o] < min_score) continue;
			keep[o] = 1;
			for (int j = match_offsets[m]; j < match_offsets[m + 1]; j++)
				if (best_obj[key_of[j]] == o) members[t++] = j - match_offsets[m];
			cluster_offsets[++ns] = t;
		}
	free(keys); free(key_of); free(best_score); free(best_obj); free(in_cl); free(owned);
	return ns;
}

/* ==== moped3d: depth-aware pose stage ====================================================
 * Restates POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CPU (moped3d/libmoped/src/pose/POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CPU.hpp);
 * pinned against the class itself compiled into oracle/_ref/libmoped3d_ref.so (ref3d_harness.cpp, tests/test_oracle3d_pose.py).
 * What differs from the moped2 stage above: every correspondence also carries world3D (the back-projected, depth-filled
 * 3-D point of the feature) and a Cauchy weight from its fill distance (:187-190, 352-353); the two residuals per
 * correspondence are the squared distances from the transformed model point to its projection on the feature's viewing ray
 * and from that projection to world3D, weighted by 1-(1-Alpha)w and (1-Alpha)w (:136-173); the initial translation is the
 * mean world3D of the samples (:262-276). randSample, testAllPoints (2-D reprojection), the RANSAC loop and levmar are the
 * same code. The stage is compiled WITHOUT -fsingle-precision-constant (moped3d/libmoped/Makefile:44-51): literals are double. */

float mo_cauchy_weight(float fill_distance) {                       /* getCauchyWeight :187-190, FillInCauchyScale = 0.100 */
	float factor = fill_distance / (float)0.100;
	return (float)(1.0 / (1 + factor * factor));
}

void mo_lm_func_depth(const float *p7, float *res, int n_pts, const float *xyz, const float *world, const float *cauchy,
                      const int *image, const mo_camera *cams, float alpha) {
	float q[4] = { p7[0], p7[1], p7[2], p7[3] };
	quat_norm(q);
	float T[12];
	tm_init(T, q, p7 + 4);
	for (int i = 0; i < n_pts; i++) {
		const mo_camera *cam = &cams[image[i]];
		float p3[3];
		tm_transform(T, p3, xyz + 3 * i);
		tm_inverse(cam->TM, p3, p3);
		if (p3[2] < 0) {
			res[2 * i] = -p3[2] + 10;
			res[2 * i + 1] = -p3[2] + 10;
		} else {
			const float *w3 = world + 3 * i;
			float vx = w3[0], vy = w3[1], vz = w3[2];
			float norm = sqrtf(vx * vx + vy * vy + vz * vz);
			float nx = vx / norm, ny = vy / norm, nz = vz / norm;
			float dot = nx * p3[0] + ny * p3[1] + nz * p3[2];
			float hx = nx * dot, hy = ny * dot, hz = nz * dot;                      /* pHat: projection of p3D on the viewing ray */
			float ax = p3[0] - hx, ay = p3[1] - hy, az = p3[2] - hz;
			float dxy = sqrtf(ax * ax + ay * ay + az * az);                        /* Pt::euclDist */
			float bx = w3[0] - hx, by = w3[1] - hy, bz = w3[2] - hz;
			float dz = sqrtf(bx * bx + by * by + bz * bz);
			res[2 * i] = dxy * dxy;
			res[2 * i + 1] = dz * dz;
		}
		float wi = cauchy[i];
		float weight3D = (1 - alpha) * wi;
		float weight2D = 1 - weight3D;
		res[2 * i] *= weight2D;
		res[2 * i + 1] *= weight3D;
	}
}

typedef struct { int n_pts; const float *xyz, *world, *cauchy; const int *image; const mo_camera *cams; float alpha; } depth_ctx;
static void depth_fn(const float *p7, float *res, const void *c) {
	const depth_ctx *d = (const depth_ctx *)c;
	mo_lm_func_depth(p7, res, d->n_pts, d->xyz, d->world, d->cauchy, d->image, d->cams, d->alpha);
}

float mo_optimize_camera_depth(float *pose7, int n_pts, int itmax, const float *xyz, const float *world, const float *cauchy,
                               const int *image, const mo_camera *cams, float alpha) {
	float p[7], info[10];
	memcpy(p, pose7, sizeof p);
	depth_ctx c = { n_pts, xyz, world, cauchy, image, cams, alpha };
	int r = levmar_dif_fn(p, 2 * n_pts, itmax, depth_fn, &c, info);
	if (r < 0) return (float)r;
	memcpy(pose7, p, sizeof p);
	quat_norm(pose7);
	return info[1];
}

/* initPose :262-276: translation = mean world3D of the samples (Pt += then / n); the quaternion is the caller's */
void mo_init_translation_depth(const float *world, const int *sample_pos, int n_samples, float *t3) {
	float sx = 0.f, sy = 0.f, sz = 0.f;
	for (int j = 0; j < n_samples; j++) { const float *w = world + 3 * sample_pos[j]; sx += w[0]; sy += w[1]; sz += w[2]; }
	t3[0] = sx / n_samples; t3[1] = sy / n_samples; t3[2] = sz / n_samples;
}

/* one RANSAC iteration body (:283-312) on explicit (sample positions, initial quaternion): like mo_hypothesis */
int mo_hypothesis_depth(int n, const float *xy, const float *xyz, const float *world, const float *cauchy, const int *image, const mo_camera *cams,
                        float alpha, const int *sample_pos, int n_samples, const float *init_quat, int max_lm, float err_thr, int min_npts,
                        float *pose_lm, float *pose_refit, float *lm_err2, unsigned char *mask) {
	float *gxyz = (float *)malloc(sizeof(float) * 3 * (size_t)(n + 1)), *gw = (float *)malloc(sizeof(float) * 3 * (size_t)(n + 1));
	float *gc = (float *)malloc(sizeof(float) * (size_t)(n + 1));
	int *gim = (int *)malloc(sizeof(int) * (size_t)(n + 1));
	for (int j = 0; j < n_samples; j++) {
		int s = sample_pos[j];
		memcpy(gxyz + 3 * j, xyz + 3 * s, 12); memcpy(gw + 3 * j, world + 3 * s, 12); gc[j] = cauchy[s]; gim[j] = image[s];
	}
	float pose[7] = { init_quat[0], init_quat[1], init_quat[2], init_quat[3], 0, 0, 0 };
	mo_init_translation_depth(world, sample_pos, n_samples, pose + 4);
	memset(mask, 0, (size_t)n);
	lm_err2[1] = -2;
	int ret = -1;
	float r = mo_optimize_camera_depth(pose, n_samples, max_lm, gxyz, gw, gc, gim, cams, alpha);
	lm_err2[0] = r;
	if ((int)r != -1) {
		memcpy(pose_lm, pose, sizeof pose);
		ret = mo_test_all_points(pose, n, xy, xyz, image, cams, err_thr, mask);
		if (ret > min_npts) {
			int k = 0;
			for (int i = 0; i < n; i++) if (mask[i]) { memcpy(gxyz + 3 * k, xyz + 3 * i, 12); memcpy(gw + 3 * k, world + 3 * i, 12); gc[k] = cauchy[i]; gim[k] = image[i]; k++; }
			lm_err2[1] = mo_optimize_camera_depth(pose, k, max_lm, gxyz, gw, gc, gim, cams, alpha);
		}
		memcpy(pose_refit, pose, sizeof pose);
	}
	free(gxyz); free(gw); free(gc); free(gim);
	return ret;
}

/* RANSAC() :278-314 with the seedable stream (same draws as ref3d_ransac: randSample then four quaternion components) */
int mo_ransac_depth(uint64_t *state, int n, const float *xy, const float *xyz, const float *world, const float *cauchy, const int *image,
                    const int *tie_ids, const mo_camera *cams, float alpha, int max_ransac, int max_lm, int n_pts_align, int min_npts,
                    float err_thr, float *pose7, int *iters) {
	int pos[16];
	float init[7], pose_lm[7], pose_refit[7], err2[2];
	unsigned char *mask = (unsigned char *)malloc((size_t)n + 1);
	int found = 0, it;
	for (it = 0; it < max_ransac; it++) {
		if (!mo_rand_sample(state, xy, image, tie_ids, n, n_pts_align, pos)) break;
		mo_init_pose(state, init);                     /* the four rand() calls of initPose; its translation is replaced below */
		int r = mo_hypothesis_depth(n, xy, xyz, world, cauchy, image, cams, alpha, pos, n_pts_align, init, max_lm, err_thr, min_npts,
		                            pose_lm, pose_refit, err2, mask);
		if (r > min_npts) { memcpy(pose7, pose_refit, sizeof pose_refit); found = 1; it++; break; }
	}
	if (iters) *iters = it;
	free(mask);
	return found;
}

/* ---- the second depth pose variant: POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CPU (moped3d/libmoped/src/pose/…:57-435) --------------
 * Three residuals per correspondence: the squared pixel differences of the moped2 stage, and 50 x the squared distance between
 * the camera-frame point p3D and (p3D . world3D) p3D — computed and stored even when the point is behind the camera (the depth
 * term sits outside the if/else, :188-200). Weights: (1 - w3D), (1 - w3D), w3D with w3D = (1 - Alpha) * cauchy, Cauchy scale 25
 * (:66). initPose, randSample, testAllPoints and RANSAC are the ones of the back-projection variant. */
float mo_cauchy_weight_v1(float fill_distance) {
	float factor = fill_distance / (float)25.0;
	return (float)(1.0 / (1 + factor * factor));
}

void mo_lm_func_depth_v1(const float *p7, float *res, int n_pts, const float *xy, const float *xyz, const float *world, const float *cauchy,
                         const int *image, const mo_camera *cams, float alpha) {
	float q[4] = { p7[0], p7[1], p7[2], p7[3] };
	quat_norm(q);
	float T[12];
	tm_init(T, q, p7 + 4);
	for (int i = 0; i < n_pts; i++) {
		const mo_camera *cam = &cams[image[i]];
		float p3[3];
		tm_transform(T, p3, xyz + 3 * i);
		tm_inverse(cam->TM, p3, p3);
		float u = p3[0] / p3[2] * cam->K[0] + cam->K[2];
		float v = p3[1] / p3[2] * cam->K[1] + cam->K[3];
		if (p3[2] < 0) {
			res[3 * i] = -p3[2] + 10;
			res[3 * i + 1] = -p3[2] + 10;
			res[3 * i + 2] = -p3[2] + 10;
		} else {
			float dx = u - xy[2 * i], dy = v - xy[2 * i + 1];
			res[3 * i] = dx * dx;
			res[3 * i + 1] = dy * dy;
		}
		const float *w3 = world + 3 * i;
		float vecTP = p3[0] * w3[0] + p3[1] * w3[1] + p3[2] * w3[2];
		float pw[3] = { p3[0] * vecTP, p3[1] * vecTP, p3[2] * vecTP };
		float dx = p3[0] - pw[0], dy = p3[1] - pw[1], dz = p3[2] - pw[2];
		float depthError = sqrtf(dx * dx + dy * dy + dz * dz);                   /* projWorld.euclDist(p3D) */
		res[3 * i + 2] = depthError * depthError;
		res[3 * i + 2] *= 50;
		float wi = cauchy[i];
		float weight3D = (1 - alpha) * wi;
		res[3 * i] *= (1 - weight3D);
		res[3 * i + 1] *= (1 - weight3D);
		res[3 * i + 2] *= weight3D;
	}
}

typedef struct { int n_pts; const float *xy, *xyz, *world, *cauchy; const int *image; const mo_camera *cams; float alpha; } depth1_ctx;
static void depth1_fn(const float *p7, float *res, const void *c) {
	const depth1_ctx *d = (const depth1_ctx *)c;
	mo_lm_func_depth_v1(p7, res, d->n_pts, d->xy, d->xyz, d->world, d->cauchy, d->image, d->cams, d->alpha);
}

static float optimize_camera_depth_v1(float *pose7, int n_pts, int itmax, const float *xy, const float *xyz, const float *world, const float *cauchy,
                                      const int *image, const mo_camera *cams, float alpha) {
	float p[7], info[10];
	memcpy(p, pose7, sizeof p);
	depth1_ctx c = { n_pts, xy, xyz, world, cauchy, image, cams, alpha };
	int r = levmar_dif_fn(p, 3 * n_pts, itmax, depth1_fn, &c, info);
	if (r < 0) return (float)r;
	memcpy(pose7, p, sizeof p);
	quat_norm(pose7);
	return info[1];
}

int mo_hypothesis_depth_v1(int n, const float *xy, const float *xyz, const float *world, const float *cauchy, const int *image, const mo_camera *cams,
                           float alpha, const int *sample_pos, int n_samples, const float *init_quat, int max_lm, float err_thr, int min_npts,
                           float *pose_lm, float *pose_refit, float *lm_err2, unsigned char *mask) {
	size_t cap = (size_t)n + 1;
	float *gxy = (float *)malloc(sizeof(float) * 2 * cap), *gxyz = (float *)malloc(sizeof(float) * 3 * cap), *gw = (float *)malloc(sizeof(float) * 3 * cap);
	float *gc = (float *)malloc(sizeof(float) * cap);
	int *gim = (int *)malloc(sizeof(int) * cap);
	for (int j = 0; j < n_samples; j++) {
		int s = sample_pos[j];
		memcpy(gxy + 2 * j, xy + 2 * s, 8); memcpy(gxyz + 3 * j, xyz + 3 * s, 12); memcpy(gw + 3 * j, world + 3 * s, 12); gc[j] = cauchy[s]; gim[j] = image[s];
	}
	float pose[7] = { init_quat[0], init_quat[1], init_quat[2], init_quat[3], 0, 0, 0 };
	mo_init_translation_depth(world, sample_pos, n_samples, pose + 4);
	memset(mask, 0, (size_t)n);
	lm_err2[1] = -2;
	int ret = -1;
	float r = optimize_camera_depth_v1(pose, n_samples, max_lm, gxy, gxyz, gw, gc, gim, cams, alpha);
	lm_err2[0] = r;
	if ((int)r != -1) {
		memcpy(pose_lm, pose, sizeof pose);
		ret = mo_test_all_points(pose, n, xy, xyz, image, cams, err_thr, mask);
		if (ret > min_npts) {
			int k = 0;
			for (int i = 0; i < n; i++) if (mask[i]) {
				memcpy(gxy + 2 * k, xy + 2 * i, 8); memcpy(gxyz + 3 * k, xyz + 3 * i, 12); memcpy(gw + 3 * k, world + 3 * i, 12); gc[k] = cauchy[i]; gim[k] = image[i]; k++;
			}
			lm_err2[1] = optimize_camera_depth_v1(pose, k, max_lm, gxy, gxyz, gw, gc, gim, cams, alpha);
		}
		memcpy(pose_refit, pose, sizeof pose);
	}
	free(gxy); free(gxyz); free(gw); free(gc); free(gim);
	return ret;
}

int mo_ransac_depth_v1(uint64_t *state, int n, const float *xy, const float *xyz, const float *world, const float *cauchy, const int *image,
                       const int *tie_ids, const mo_camera *cams, float alpha, int max_ransac, int max_lm, int n_pts_align, int min_npts,
                       float err_thr, float *pose7, int *iters) {
	int pos[16];
	float init[7], pose_lm[7], pose_refit[7], err2[2];
	unsigned char *mask = (unsigned char *)malloc((size_t)n + 1);
	int found = 0, it;
	for (it = 0; it < max_ransac; it++) {
		if (!mo_rand_sample(state, xy, image, tie_ids, n, n_pts_align, pos)) break;
		mo_init_pose(state, init);
		int r = mo_hypothesis_depth_v1(n, xy, xyz, world, cauchy, image, cams, alpha, pos, n_pts_align, init, max_lm, err_thr, min_npts,
		                               pose_lm, pose_refit, err2, mask);
		if (r > min_npts) { memcpy(pose7, pose_refit, sizeof pose_refit); found = 1; it++; break; }
	}
	if (iters) *iters = it;
	free(mask);
	return found;
}
