/*
 * moped_linkage_oracle.c — CPU restatement of moped3d's clustering stage CLUSTER_LINKAGE_CPU
 * (moped3d/libmoped/src/cluster/CLUSTER_LINKAGE_CPU.hpp:49-706). TEST INFRASTRUCTURE ONLY (see
 * moped_oracle.h); there is no CUDA kernel for this stage yet — this file and its pinning tests are the checker it will be built
 * against. Pinned by tests/test_oracle3d_linkage.py against the class itself compiled unmodified into
 * oracle/_ref/libmoped3d_ref_strict.so (bit for bit) and oracle/_ref/libmoped3d_ref.so (the reference's own -ffast-math flags).
 *
 * The stage, per model: pairwise similarity matrices over the model's matches — Gaussian kernels on image distance (K2D) and
 * on the distance of the back-projected 3-D points (K3D) with bandwidths = average nearest-neighbour distances (:98-123,
 * 133-150), a depth-discontinuity kernel from the steepest slope change along the Bresenham path between two features in the
 * depth map (:167-285), a model-vs-world distance consistency kernel (:152-172) — combined (sum or product, each followed by a
 * division by the maximum, :287-322, 624-648), blended with K2D by per-feature Cauchy weights of the depth fill distance
 * (:324-365), then agglomerative clustering with minimum / average / maximum linkage down to a similarity cutoff (:414-531).
 * Quirks kept because they decide the output: the merged-away cluster index stays in the candidate list until the scan reaches
 * it and the element after it is skipped in that scan (:449-455); `valid[]` is never cleared (:519-527); adaptiveWeightSum is
 * called with alpha = 0.5, gamma = 25 whatever the constructor got (:651); literals are double (the stage is compiled without
 * -fsingle-precision-constant).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "moped_oracle.h"

#define AT(K, N, x, y) ((K)[(size_t)(y) * (N) + (x)])            /* Image::getProb/setProb(x, y): row-major, x fastest */

static float sq_dist(const float *a, const float *b, int dim) {   /* Pt::sqEuclDist, moped.hpp:125 */
	float d, r = 0;
	for (int x = 0; x < dim; x++) { d = b[x] - a[x]; r += d * d; }
	return r;
}

static void avg_nn_distances(int n, const float *xy, const float *xyz, float *nn2D, float *nn3D) {   /* :98-123 */
	float s2 = 0, s3 = 0;
	for (int i = 0; i < n; i++) {
		float b2 = (float)DBL_MAX, b3 = (float)DBL_MAX;
		for (int j = 0; j < n; j++) {
			if (i == j) continue;
			float d2 = sqrtf(sq_dist(xy + 2 * j, xy + 2 * i, 2)), d3 = sqrtf(sq_dist(xyz + 3 * j, xyz + 3 * i, 3));
			if (b2 > d2) b2 = d2;
			if (b3 > d3) b3 = d3;
		}
		s2 += b2; s3 += b3;
	}
	*nn2D = s2 / n; *nn3D = s3 / n;
}

static void gauss_k(float *K, int n, const float *pts, int dim, float sigma) {                       /* :133-150 */
	float two = 2 * sigma * sigma;
	for (int i = 0; i < n; i++)
		for (int j = i; j < n; j++) {
			float distance = sq_dist(pts + (size_t)dim * i, pts + (size_t)dim * j, dim);
			float val = expf(-1 * distance / two);
			AT(K, n, i, j) = val; AT(K, n, j, i) = val;
		}
}

static void filter3d_k(float *K, int n, const float *xyz, const float *world) {                       /* :152-172 */
	float sigma = (float)0.1;
	float two = 2 * sigma * sigma;
	for (int i = 0; i < n; i++) {
		AT(K, n, i, i) = (float)1.0;
		for (int j = i + 1; j < n; j++) {
			float dm = sqrtf(sq_dist(xyz + 3 * j, xyz + 3 * i, 3));
			float dw = sqrtf(sq_dist(world + 3 * j, world + 3 * i, 3));
			float e = fabsf(dm - dw) / dm;
			float val = expf((-1 * e * e) / two);
			AT(K, n, i, j) = val; AT(K, n, j, i) = val;
		}
	}
}

typedef struct { int x, y; } ipair;

/* bresenhamIterate :175-223; returns the number of coordinates written */
static int bresenham(ipair *out, ipair p, ipair q, int n_samples) {
	int x0 = p.x, y0 = p.y, x1 = q.x, y1 = q.y, t, cnt = 0;
	int steep = abs(y1 - y0) > abs(x1 - x0);
	if (steep) { t = x0; x0 = y0; y0 = t; t = x1; x1 = y1; y1 = t; }
	if (x0 > x1) { t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
	float deltaX = (float)x1 - x0, deltaY = fabsf((float)y1 - y0);
	int yStep = (y0 < y1) ? 1 : -1;
	int perStep = (int)(x1 - x0) / n_samples;
	if (perStep < 1) perStep = 1;
	float error = 0.0, deltaError = ((float)deltaY) / deltaX;
	float intPart;
	int y = y0;
	for (int x = x0; x <= x1;) {
		if (steep) { out[cnt].x = y; out[cnt].y = x; } else { out[cnt].x = x; out[cnt].y = y; }
		cnt++;
		x += perStep;
		if (x > x1) break;                          /* the reference still updates error/y here; the values are never used */
		error += deltaError * perStep * yStep;
		error = modff(error, &intPart);
		y += intPart;
	}
	return cnt;
}

static ipair saturate(int x, int y, int W, int H) {                                                   /* :225-231 */
	ipair r;
	x = (x < 0) ? 0 : x; x = (x >= W) ? W - 1 : x;
	y = (y < 0) ? 0 : y; y = (y >= H) ? H - 1 : y;
	r.x = x; r.y = y;
	return r;
}

static void discontinuity_k(float *K, int n, const float *xy, int W, int H, const float *depth) {     /* :233-285 */
	ipair *coords = (ipair *)malloc(sizeof(ipair) * (size_t)(W + H + 4));
	float div = (float)(-2 * (M_PI / 128) * (M_PI / 128));
	for (int i = 0; i < n; i++) {
		ipair li = saturate((int)xy[2 * i], (int)xy[2 * i + 1], W, H);
		for (int j = i; j < n; j++) {
			ipair lj = saturate((int)xy[2 * j], (int)xy[2 * j + 1], W, H);
			int cnt = bresenham(coords, li, lj, 20);
			float dStart = depth[(size_t)li.y * W + li.x], dEnd = depth[(size_t)lj.y * W + lj.x];
			int xd = li.x - lj.x, yd = li.y - lj.y;
			float planeDist = sqrtf((float)(xd * xd + yd * yd));
			float direct = atan2f(dEnd - dStart, planeDist);
			float maxDiff = -1;
			for (int k = 0; k < cnt - 1; k++) {
				ipair c1 = coords[k], c2 = coords[k + 1];
				float d1 = depth[(size_t)c1.y * W + c1.x], d2 = depth[(size_t)c2.y * W + c2.x];
				float dx = c1.x - c2.x, dy = c1.y - c2.y;
				float pixDist = sqrtf(dx * dx + dy * dy);
				float pixAngle = atan2f(d2 - d1, pixDist);
				float diff = fabsf(direct - pixAngle);
				if (diff > maxDiff) maxDiff = diff;
			}
			float val = expf(maxDiff * maxDiff / div);
			AT(K, n, i, j) = val; AT(K, n, j, i) = val;
		}
	}
	free(coords);
}

static void normalize_k(float *K, int n) {                                                            /* :303-322 */
	float mx = -1;
	for (size_t k = 0; k < (size_t)n * n; k++) if (K[k] > mx) mx = K[k];
	for (size_t k = 0; k < (size_t)n * n; k++) K[k] = K[k] / mx;
}

/* adaptiveWeightSum :324-365 */
static void adaptive_weight_sum(float *K, int n, const float *xy, int W, const float *distance, const float *K2D, const float *K3D,
                                float alpha, float gamma) {
	float gammaSq = gamma * gamma;
	float *w = (float *)malloc(sizeof(float) * (size_t)(n + 1));
	for (int i = 0; i < n; i++) {
		float d = distance[(size_t)((int)xy[2 * i + 1]) * W + (int)xy[2 * i]];
		w[i] = (float)(1.0 / (1 + (d * d / gammaSq)));
	}
	float alphaBar = (float)(1.0 - alpha);
	for (int i = 0; i < n; i++)
		for (int j = i; j < n; j++) {
			float k2 = AT(K2D, n, i, j), k3 = AT(K3D, n, i, j);
			float joint = w[i] * w[j];
			float w2D = (float)(alpha + alphaBar * (1.0 - joint)), w3D = alphaBar * joint;
			float val = w2D * k2 + w3D * k3;
			AT(K, n, i, j) = val; AT(K, n, j, i) = val;
		}
	free(w);
}

/* hierarchicalCluster :414-531. Clusters are index lists; a merge appends the second cluster back to front. */
static int hierarchical_cluster(const float *K, int n, float cutoff, int min_pts, int linkage, int *cluster_offsets, int *members) {
	int **cl = (int **)malloc(sizeof(int *) * (size_t)(n + 1));
	int *sz = (int *)calloc((size_t)n + 1, sizeof(int)), *cap = (int *)calloc((size_t)n + 1, sizeof(int));
	for (int i = 0; i < n; i++) { cl[i] = (int *)malloc(sizeof(int) * 4); cap[i] = 4; cl[i][0] = i; sz[i] = 1; }
	float *D = (float *)malloc(sizeof(float) * (size_t)n * n + 4);
	int *valid = (int *)malloc(sizeof(int) * (size_t)(n + 1)), nv = n;      /* validIndices: an ordered list with erase */
	for (int i = 0; i < n; i++) {
		valid[i] = i;
		for (int j = i; j < n; j++) D[(size_t)j * n + i] = D[(size_t)i * n + j] = AT(K, n, i, j);
	}
	int removeValue = -1;
	for (;;) {
		float maxSim = -1;
		int p1 = 0, p2 = 0;
		for (int a = 0; a < nv; a++) {
			int index1 = valid[a];
			if (index1 == removeValue) {
				/* `index1_it = validIndices.erase(index1_it); continue;` — the for's ++ then skips the element that followed */
				memmove(valid + a, valid + a + 1, sizeof(int) * (size_t)(nv - a - 1));
				nv--;
				continue;
			}
			for (int b = a + 1; b < nv; b++) {
				int index2 = valid[b];
				if (D[(size_t)index1 * n + index2] > maxSim) { maxSim = D[(size_t)index1 * n + index2]; p1 = index1; p2 = index2; }
			}
		}
		if (maxSim < cutoff) break;
		int sU = sz[p1], sR = sz[p2];
		while (sz[p2] != 0) {
			if (sz[p1] == cap[p1]) { cap[p1] *= 2; cl[p1] = (int *)realloc(cl[p1], sizeof(int) * (size_t)cap[p1]); }
			cl[p1][sz[p1]++] = cl[p2][--sz[p2]];
		}
		int toUpdate = p1;
		removeValue = p2;
		for (int i = 0; i < n; i++) {
			if (linkage == 1) {
				D[(size_t)toUpdate * n + i] = (float)((1.0 / (sU + sR)) * (sU * D[(size_t)toUpdate * n + i] + sR * D[(size_t)removeValue * n + i]));
				D[(size_t)i * n + toUpdate] = D[(size_t)toUpdate * n + i];
			} else {
				/* minimumLinkage / maximumLinkage over the CURRENT member lists of clusters i and toUpdate (:383-411) */
				float v = linkage == 0 ? (float)1e20 : -1;
				for (int a = 0; a < sz[i]; a++)
					for (int b = 0; b < sz[toUpdate]; b++) {
						float k = AT(K, n, cl[i][a], cl[toUpdate][b]);
						if (linkage == 0 ? (k < v) : (k > v)) v = k;
					}
				D[(size_t)toUpdate * n + i] = D[(size_t)i * n + toUpdate] = v;
			}
		}
	}
	int nc = 0, k = 0;
	cluster_offsets[0] = 0;
	for (int i = 0; i < n; i++) {
		if (sz[i] > min_pts) {
			memcpy(members + k, cl[i], sizeof(int) * (size_t)sz[i]);
			k += sz[i];
			cluster_offsets[++nc] = k;
		}
		free(cl[i]);
	}
	free(cl); free(sz); free(cap); free(D); free(valid);
	return nc;
}

/* The similarity matrix K of CLUSTER_LINKAGE_CPU::process for ONE model's matches (:596-651); K is n x n, symmetric. */
void mo_linkage_similarity(int n, const float *xy, const float *xyz, const float *world, int W, int H, const float *depth, const float *distance,
                           int use3DFilter, float sigma2D, float sigma3D, float *K) {
	if (n <= 0) return;
	float k2s, k3s;
	if (sigma2D == -1 || sigma3D == -1) {
		avg_nn_distances(n, xy, xyz, &k2s, &k3s);
		if (sigma2D != -1) k2s = sigma2D;
		if (sigma3D != -1) k3s = sigma3D;
	} else { k2s = sigma2D; k3s = sigma3D; }
	size_t nn = (size_t)n * n;
	float *K2D = (float *)malloc(sizeof(float) * nn), *K3D = (float *)malloc(sizeof(float) * nn), *BK = (float *)malloc(sizeof(float) * nn);
	gauss_k(K2D, n, xy, 2, k2s);
	gauss_k(K3D, n, world, 3, k3s);
	discontinuity_k(BK, n, xy, W, H, depth);
	for (size_t k = 0; k < nn; k++) K3D[k] = K3D[k] + BK[k];                 /* getSum(K3D, BK, K3D) */
	normalize_k(K3D, n);
	if (use3DFilter) {
		float *K3F = (float *)malloc(sizeof(float) * nn);
		filter3d_k(K3F, n, xyz, world);
		if (use3DFilter == 1) for (size_t k = 0; k < nn; k++) K3D[k] = K3D[k] + K3F[k];
		else for (size_t k = 0; k < nn; k++) K3D[k] = K3D[k] * K3F[k];
		normalize_k(K3D, n);
		free(K3F);
	}
	adaptive_weight_sum(K, n, xy, W, distance, K2D, K3D, (float)0.5, 25);
	free(K2D); free(K3D); free(BK);
}

/* hierarchicalCluster (:414-531) on a given similarity matrix */
int mo_linkage_agglomerate(const float *K, int n, float cutoff, int min_pts, int linkage_type, int *cluster_offsets, int *members) {
	cluster_offsets[0] = 0;
	if (n <= 0) return 0;
	return hierarchical_cluster(K, n, cutoff, min_pts, linkage_type, cluster_offsets, members);
}

/* CLUSTER_LINKAGE_CPU::process for ONE model's matches (:596-700). depth / distance: W x H row-major maps. */
int mo_cluster_linkage(int n, const float *xy, const float *xyz, const float *world, int W, int H, const float *depth, const float *distance,
                       float cutoff, int min_pts, int use3DFilter, int linkage_type, float sigma2D, float sigma3D,
                       int *cluster_offsets, int *members) {
	cluster_offsets[0] = 0;
	if (n <= 0) return 0;
	float *K = (float *)malloc(sizeof(float) * (size_t)n * n);
	mo_linkage_similarity(n, xy, xyz, world, W, H, depth, distance, use3DFilter, sigma2D, sigma3D, K);
	int nc = hierarchical_cluster(K, n, cutoff, min_pts, linkage_type, cluster_offsets, members);
	free(K);
	return nc;
}
