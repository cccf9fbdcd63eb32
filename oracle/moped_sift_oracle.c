/*
 * moped_sift_oracle.c — CPU restatement of MOPED's step 1 (feature extraction): FEAT_SIFT_CPU over the vendored
 * libsiftfast 1.1. TEST INFRASTRUCTURE ONLY (see moped_oracle.h): the checker of the CUDA extractor
 * (moped_b200/csrc/sift.cu), never linked into the product.
 *
 * What is restated (file:line relative to /root/reference; libs.tgz! = moped2/libmoped/libs/libs.tgz):
 *   FEAT_SIFT_CPU::process           moped2/libmoped/src/feat/FEAT_SIFT_CPU.hpp:78-112
 *   GetKeypoints .. PlaceInIndex     libs.tgz!libsiftfast-1.1-src/libsiftfast.cpp:301-1668
 * The vendored copy #undef's __SSE__/__SSE2__/__SSE3__ (libsiftfast.cpp:40-42), so the reference runs the SCALAR
 * branches: ConvHorizontal/ConvVertical/ConvBuffer (:523-581), GradOriImages with libm atan2f (:959-992), the
 * scalar descriptor normalisation (:1503-1516). Those are what this file follows.
 *
 * Arithmetic: every sum is written in the reference's source order. The convolution taps (ConvBuffer) accumulate with
 * fused multiply-add in tap order; everything else uses separate multiply and add (this file is compiled with
 * -ffp-contract=off). The reference itself is built with -ffast-math -march=native: its compiler may re-associate,
 * vectorise and contract the sums (the x86-64-v3 build in oracle/_ref has 8-lane partial sums and a mix of vfmadd and
 * vmul/vadd in ConvBuffer), so there is no single "reference rounding"; both variants of this file were measured to be
 * equally close to oracle/_ref (same keypoint lists, mean |d coord2D| 3e-5 px). Parity against oracle/_ref is therefore
 * to tolerance (tests/test_sift_oracle.py states it); parity CUDA-vs-this-file is exact up to the libm calls (expf,
 * atan2f, sinf, cosf, powf).
 *
 * Keypoint order: the reference prepends to a linked list (libsiftfast.cpp:941-951,1423); with one OpenMP thread
 * the list FEAT_SIFT_CPU walks is the exact reverse of creation order. With several threads the row order and the
 * duplicate suppression (:1188-1196) race; the single-thread order is the one restated here.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "moped_oracle.h"

#define SIFT_PI 3.141592654f
#define SIFT_SQRT2 1.4142136f
#define SIFT_SCALES 3
#define SIFT_INIT_SIGMA 1.6f

typedef struct { int rows, cols; float *px; } simg;

static simg img_new(int rows, int cols) {
	simg m; m.rows = rows; m.cols = cols;
	m.px = (float *)calloc((size_t)rows * cols + 8, sizeof(float));
	return m;
}

/* GaussianBlur's kernel (libsiftfast.cpp:470-508): ksize+1 weights are summed, ksize are normalised. */
int mo_sift_gauss_kernel(float fblur, float *kernel /* >= 64 */) {
	const float GaussTruncate = 4.0f;
	int ksize = (int)(2.0f * GaussTruncate * fblur + 1.0f);
	if (ksize < 3) ksize = 3;
	ksize += !(ksize & 1);
	double faccum = 0;
	int width = ksize >> 1;
	for (int i = 0; i <= ksize; ++i) {
		float fweight = expf(-(float)(i - width) * (i - width) / (2.0f * fblur * fblur));
		faccum += (double)fweight;
		kernel[i] = fweight;
	}
	for (int i = 0; i < ksize; ++i) kernel[i] /= (float)faccum;
	return ksize;
}

/* How the convolution taps accumulate: 1 (default) = fused multiply-add in tap order — the variant the CUDA kernels
 * implement; 0 = separate multiply and add in tap order — what a strict-IEEE build of libsiftfast.cpp computes, used by
 * tests/test_oracle_vs_strict_ref.py to compare this whole file with such a build BIT FOR BIT. */
static int g_conv_fma = 1;
void mo_sift_set_conv_fma(int on) { g_conv_fma = on != 0; }

/* ConvBuffer (:573-581) on a replicate-padded line */
static void conv_line(const float *buf, const float *kernel, int n, int ksize, float *out, int ostride) {
	for (int i = 0; i < n; ++i) {
		float faccum = 0;
		if (g_conv_fma) for (int j = 0; j < ksize; ++j) faccum = fmaf(buf[i + j], kernel[j], faccum);
		else for (int j = 0; j < ksize; ++j) faccum += buf[i + j] * kernel[j];
		out[(size_t)i * ostride] = faccum;
	}
}

/* GaussianBlur (:470-521) = ConvHorizontal (:523-546) src->dst, then ConvVertical (:548-571) in place */
static void gaussian_blur(simg dst, simg src, float fblur) {
	float kernel[80];
	int ksize = mo_sift_gauss_kernel(fblur, kernel);
	int width = ksize >> 1, rows = src.rows, cols = src.cols;
	int n = (rows > cols ? rows : cols) + ksize;
	float *buf = (float *)malloc(sizeof(float) * n);
	for (int i = 0; i < rows; ++i) {
		const float *p = src.px + (size_t)i * cols;
		for (int j = 0; j < width; ++j) buf[j] = p[0];
		for (int j = 0; j < cols; ++j) buf[width + j] = p[j];
		for (int j = 0; j < width; ++j) buf[cols + width + j] = p[cols - 1];
		conv_line(buf, kernel, cols, ksize, dst.px + (size_t)i * cols, 1);
	}
	for (int j = 0; j < cols; ++j) {
		float *p = dst.px + j;
		for (int i = 0; i < width; ++i) buf[i] = p[0];
		for (int i = 0; i < rows; ++i) buf[width + i] = p[(size_t)i * cols];
		for (int i = 0; i < width; ++i) buf[rows + width + i] = p[(size_t)(rows - 1) * cols];
		conv_line(buf, kernel, rows, ksize, p, cols);
	}
	free(buf);
}

/* GradOriImages (:959-992) */
static void grad_ori(simg im, simg grad, simg ori) {
	int rows = im.rows, cols = im.cols;
	for (int i = 0; i < rows; ++i) {
		const float *p = im.px + (size_t)i * cols;
		for (int j = 0; j < cols; ++j) {
			float fdiffc, fdiffr;
			if (j == 0) fdiffc = 2.0f * (p[1] - p[0]);
			else if (j == cols - 1) fdiffc = 2.0f * (p[j] - p[j - 1]);
			else fdiffc = p[j + 1] - p[j - 1];
			if (i == 0) fdiffr = 2.0f * (p[j] - p[cols + j]);
			else if (i == rows - 1) fdiffr = 2.0f * (p[-cols + j] - p[j]);
			else fdiffr = p[-cols + j] - p[cols + j];
			grad.px[(size_t)i * cols + j] = sqrtf(fdiffc * fdiffc + fdiffr * fdiffr);
			ori.px[(size_t)i * cols + j] = atan2f(fdiffr, fdiffc);
		}
	}
}

/* LocalMaxMin (:1126-1147) */
static int local_max_min(float fval, simg d, int r, int c) {
	for (int row = r - 1; row <= r + 1; ++row) {
		const float *pf = d.px + (size_t)row * d.cols + c - 1;
		if (fval > 0) { if (pf[0] > fval || pf[1] > fval || pf[2] > fval) return 0; }
		else { if (fval > pf[0] || fval > pf[1] || fval > pf[2]) return 0; }
	}
	return 1;
}

/* NotOnEdge (:1149-1162) */
static int not_on_edge(simg d, int row, int col) {
	int s = d.cols;
	const float *p = d.px + (size_t)row * s;
	float f1 = p[-s + col] - p[col] * 2 + p[s + col];
	float f2 = p[col - 1] - p[col] * 2 + p[col + 1];
	float f3 = p[s + col + 1] - p[s + col - 1];
	float f4 = p[-s + col + 1] - p[-s + col - 1];
	float f5 = (f3 - f4) * 0.25f;
	float f6 = f1 * f2 - f5 * f5;
	float f8 = f1 + f2;
	return f6 * 11 * 11 > f8 * f8 * 10;
}

/* SolveLinearSystem (:1235-1274), dim = 3 */
static void solve3(float *Y, float *H) {
	const int dim = 3;
	int bestj = 0;
	for (int i = 0; i < dim - 1; ++i) {
		float fmax = -1;
		for (int j = i; j < dim; ++j) {
			float f = H[j * dim + i];
			if (f < 0) f = -f;
			if (f > fmax) { fmax = f; bestj = j; }
		}
		if (bestj != i) {
			for (int j = 0; j < dim; ++j) { float t = H[bestj * dim + j]; H[bestj * dim + j] = H[i * dim + j]; H[i * dim + j] = t; }
			float t = Y[bestj]; Y[bestj] = Y[i]; Y[i] = t;
		}
		for (int j = i + 1; j < dim; ++j) {
			float f = H[j * dim + i] / H[i * dim + i];
			for (int k = i; k < dim; ++k) H[j * dim + k] -= f * H[i * dim + k];
			Y[j] -= Y[i] * f;
		}
	}
	for (int i = dim - 1; i >= 0; --i) {
		for (int j = dim - 1; j > i; --j) Y[i] -= Y[j] * H[i * dim + j];
		Y[i] /= H[i * dim + i];
	}
}

/* FitQuadratic (:1208-1231) */
static float fit_quadratic(float *X, const simg *dog, int index, int r, int c) {
	float H[9], Y[3];
	int s = dog[index].cols;
	const float *p0 = dog[index - 1].px + (size_t)r * s;
	const float *p1 = dog[index].px + (size_t)r * s;
	const float *p2 = dog[index + 1].px + (size_t)r * s;
	Y[0] = 0.5f * (p2[c] - p0[c]);
	Y[1] = 0.5f * (p1[s + c] - p1[-s + c]);
	Y[2] = 0.5f * (p1[c + 1] - p1[c - 1]);
	H[0] = p0[c] - 2.0f * p1[c] + p2[c];
	H[4] = p1[-s + c] - 2.0f * p1[c] + p1[s + c];
	H[8] = p1[c - 1] - 2.0f * p1[c] + p1[c + 1];
	H[3] = H[1] = 0.25f * ((p2[s + c] - p2[-s + c]) - (p0[s + c] - p0[-s + c]));
	H[6] = H[2] = 0.25f * ((p2[c + 1] - p2[c - 1]) - (p0[c + 1] - p0[c - 1]));
	H[7] = H[5] = 0.25f * ((p1[s + c + 1] - p1[s + c - 1]) - (p1[-s + c + 1] - p1[-s + c - 1]));
	X[0] = -Y[0]; X[1] = -Y[1]; X[2] = -Y[2];
	solve3(X, H);
	return p1[c] + 0.5f * (X[0] * Y[0] + X[1] * Y[1] + X[2] * Y[2]);
}

typedef struct {
	float *xy, *so, *desc;      /* creation order */
	int n, cap;
	mo_sift_trace *trace; int n_trace, cap_trace, octave;
} kplist;

static void kp_push(kplist *L, float col, float row, float scale, float ori, const float *desc) {
	if (L->n == L->cap) {
		L->cap = L->cap ? 2 * L->cap : 1024;
		L->xy = (float *)realloc(L->xy, sizeof(float) * 2 * L->cap);
		L->so = (float *)realloc(L->so, sizeof(float) * 2 * L->cap);
		L->desc = (float *)realloc(L->desc, sizeof(float) * 128 * L->cap);
	}
	L->xy[2 * L->n] = col; L->xy[2 * L->n + 1] = row;
	L->so[2 * L->n] = scale; L->so[2 * L->n + 1] = ori;
	memcpy(L->desc + (size_t)128 * L->n, desc, sizeof(float) * 128);
	L->n++;
}

/* PlaceInIndex (:1609-1668) */
static void place_in_index(float *fdesc, float mag, float ori, float rx, float cx) {
	float oribin = ori * (8.0f / (2 * (float)SIFT_PI));
	int newrow, newcol, neworient;
	float rfrac, cfrac, ofrac;
	newrow = rx < 0 ? (int)(rx - 1) : (int)rx;
	rfrac = rx - (float)newrow;
	newcol = cx < 0 ? (int)(cx - 1) : (int)cx;
	cfrac = cx - (float)newcol;
	neworient = oribin < 0 ? (int)(oribin - 1) : (int)oribin;
	ofrac = oribin - (float)neworient;
	for (int i = 0; i < 2; ++i) {
		if ((unsigned)(i + newrow) >= 4) continue;
		float frowgrad = i == 0 ? mag * (1 - rfrac) : mag * rfrac;
		for (int j = 0; j < 2; ++j) {
			if ((unsigned)(j + newcol) >= 4) continue;
			float fcolgrad = j == 0 ? frowgrad * (1 - cfrac) : frowgrad * cfrac;
			float *pf = fdesc + 8 * (4 * (i + newrow) + j + newcol);
			for (int k = 0; k < 2; ++k) {
				float forigrad = k == 0 ? fcolgrad * (1 - ofrac) : fcolgrad * ofrac;
				pf[(neworient + k) & 7] += forigrad;
			}
		}
	}
}

/* NormalizeVec (:1519-1528) */
static void normalize_vec(float *pf, int num) {
	float faccum = 0;
	for (int i = 0; i < num; ++i) faccum += pf[i] * pf[i];
	faccum = 1 / sqrtf(faccum);
	for (int i = 0; i < num; ++i) pf[i] *= faccum;
}

/* MakeKeypoint (:1409-1431) + MakeKeypointSample scalar branch (:1433-1438,1503-1516) + KeySample/AddSample (:1530-1607) */
static void make_keypoint(kplist *L, simg grad, simg orim, float fscale, float fSize, float frowstart, float fcolstart, float forient) {
	float fdesc[128];
	memset(fdesc, 0, sizeof(fdesc));
	int rows = grad.rows, cols = grad.cols;
	int rowstart = (int)(frowstart + 0.5f);
	int colstart = (int)(fcolstart + 0.5f);
	float sinang = sinf(forient), cosang = cosf(forient);
	float fdrow = frowstart - (float)rowstart;
	float fdcol = fcolstart - (float)colstart;
	float frealsize = 3.0f * fSize;
	float firealsize = 1.0f / (3.0f * fSize);
	int windowsize = (int)(frealsize * SIFT_SQRT2 * 5.0f * 0.5f + 0.5f);
	float fsr = sinang * firealsize, fcr = cosang * firealsize, fdrr = -fdrow * firealsize, fdcr = -fdcol * firealsize;
	for (int row = -windowsize; row <= windowsize; ++row) {
		float frow = (float)row;
		float fcol = -(float)windowsize;
		for (int col = -windowsize; col <= windowsize; ++col, fcol += 1) {
			float rpos = fsr * fcol + fcr * frow + fdrr;
			float cpos = fcr * fcol - fsr * frow + fdcr;
			float rx = rpos + (2.0f - 0.5f);
			float cx = cpos + (2.0f - 0.5f);
			if (rx > -0.9999f && rx < 3.9999f && cx > -0.9999f && cx < 3.9999f) {
				int r = rowstart + row, c = colstart + col;
				if (r < 0 || r >= rows || c < 0 || c >= cols) continue;
				float fgrad = grad.px[(size_t)r * cols + c] * expf(-0.125f * (rpos * rpos + cpos * cpos));
				float fo = orim.px[(size_t)r * cols + c] - forient;
				while (fo > 2 * SIFT_PI) fo -= 2 * SIFT_PI;
				while (fo < 0) fo += 2 * SIFT_PI;
				place_in_index(fdesc, fgrad, fo, rx, cx);
			}
		}
	}
	normalize_vec(fdesc, 128);
	int brenormalize = 0;
	for (int i = 0; i < 128; ++i) if (fdesc[i] > 0.2f) { fdesc[i] = 0.2f; brenormalize = 1; }
	if (brenormalize) normalize_vec(fdesc, 128);
	kp_push(L, fscale * fcolstart, fscale * frowstart, fscale * fSize, forient, fdesc);
}

/* SmoothHistogram (:1395-1407) */
static void smooth_histogram(float *phist, int numbins) {
	float ffirst = phist[0];
	float fprev = phist[numbins - 1];
	for (int i = 0; i < numbins - 1; ++i) {
		float forg = phist[i];
		phist[i] = (fprev + forg + phist[i + 1]) * 0.33333333f;
		fprev = forg;
	}
	phist[numbins - 1] = (fprev + phist[numbins - 1] + ffirst) * 0.3333333f;
}

/* AssignOriHist (:1276-1382), scalar maximum (:1352-1356), InterpPeak (:1384-1393) */
static void assign_ori_hist(kplist *L, simg grad, simg orim, float fscale, float fSize, float frowstart, float fcolstart) {
	int rowstart = (int)(frowstart + 0.5f);
	int colstart = (int)(fcolstart + 0.5f);
	int rows = grad.rows, cols = grad.cols;
	float hists[36];
	float fexpmult = -1.0f / (2.0f * 1.5f * 1.5f * fSize * fSize);
	memset(hists, 0, sizeof(hists));
	const float fbinmult = 36.0f / (2 * SIFT_PI);
	const float fbinadd = (float)(SIFT_PI + 0.001f) * fbinmult;
	int windowsize = (int)(fSize * 1.5f * 3.0f);
	for (int rowcur = rowstart - windowsize; rowcur <= rowstart + windowsize; ++rowcur) {
		if (rowcur < 0 || rowcur >= rows - 2) continue;
		for (int colcur = colstart - windowsize; colcur <= colstart + windowsize; ++colcur) {
			if (colcur < 0 || colcur >= cols - 2) continue;
			float fdx = grad.px[(size_t)rowcur * cols + colcur];
			if (fdx > 0) {
				float fdrow = (float)rowcur - frowstart, fdcol = (float)colcur - fcolstart;
				float fradius2 = fdrow * fdrow + fdcol * fdcol;
				if ((float)(windowsize * windowsize) + 0.5f > fradius2) {
					float fweight = expf(fradius2 * fexpmult);
					int binindex = (int)(orim.px[(size_t)rowcur * cols + colcur] * fbinmult + fbinadd);
					if (binindex > 36) binindex = 0;
					if (binindex == 36) binindex = 35;
					hists[binindex] += fdx * fweight;
				}
			}
		}
	}
	for (int i = 0; i < 6; ++i) smooth_histogram(hists, 36);
	float fmaxval = 0;
	for (int i = 0; i < 36; ++i) if (hists[i] > fmaxval) fmaxval = hists[i];
	fmaxval *= 0.8f;
	const float foriadd = 0.5f * 2 * SIFT_PI / 36.0f - SIFT_PI, forimult = 2 * SIFT_PI / 36.0f;
	int previndex = 35;
	for (int index = 0; index < 36; ++index) {
		if (index != 0) previndex = index - 1;
		int nextindex = 0;
		if (index != 35) nextindex = index + 1;
		if (hists[index] <= hists[previndex] || hists[index] <= hists[nextindex] || hists[index] < fmaxval) continue;
		float f0 = hists[previndex], f1 = hists[index], f2 = hists[nextindex];
		if (f1 < 0) { f0 = -f0; f1 = -f1; f2 = -f2; }
		float fpeak = 0.5f * (f0 - f2) / (f0 - 2.0f * f1 + f2);
		float forient = (index + fpeak) * forimult + foriadd;
		make_keypoint(L, grad, orim, fscale, fSize, frowstart, fcolstart, forient);
	}
}

/* InterpKeyPoint (:1164-1205); the recursion is a loop here */
static void interp_keypoint(kplist *L, const simg *dog, int index, int rowstart, int colstart, simg grad, simg orim,
                            char *maxmin, float fscale, float peak_thresh, int scan_r, int scan_c) {
	float X[3], fquad;
	int steps = 5;
	for (;;) {
		fquad = fit_quadratic(X, dog, index, rowstart, colstart);
		int newrow = rowstart, newcol = colstart;
		if (X[1] > 0.6f && rowstart < dog[0].rows - 3) newrow++;
		if (X[1] < -0.6f && rowstart > 3) newrow--;
		if (X[2] > 0.6f && colstart < dog[0].cols - 3) newcol++;
		if (X[2] < -0.6f && colstart > 3) newcol--;
		if (steps > 0 && (newrow != rowstart || newcol != colstart)) { rowstart = newrow; colstart = newcol; steps--; continue; }
		break;
	}
	if (fabsf(X[0]) <= 1.5f && fabsf(X[1]) <= 1.5f && fabsf(X[2]) <= 1.5f && fabsf(fquad) >= peak_thresh) {
		char *pm = maxmin + (size_t)rowstart * grad.cols + colstart;
		if (!pm[0]) {
			pm[0] = 1;
			float fSize = SIFT_INIT_SIGMA * powf(2.0f, ((float)index + X[0]) / (float)SIFT_SCALES);
			if (L->trace) {
				if (L->n_trace == L->cap_trace) {
					L->cap_trace = L->cap_trace ? 2 * L->cap_trace : 1024;
					L->trace = (mo_sift_trace *)realloc(L->trace, sizeof(mo_sift_trace) * L->cap_trace);
				}
				mo_sift_trace *t = &L->trace[L->n_trace++];
				t->octave = L->octave; t->index = index; t->scan_row = scan_r; t->scan_col = scan_c;
				t->row = rowstart; t->col = colstart; t->X[0] = X[0]; t->X[1] = X[1]; t->X[2] = X[2]; t->fsize = fSize;
				t->first_kp = L->n;
			}
			assign_ori_hist(L, grad, orim, fscale, fSize, (float)rowstart + X[1], (float)colstart + X[2]);
		}
	}
}

/* GetKeypoints (:301-361) with OctaveKeypoints (:410-437), FindMaxMin (:891-957), SiftDoubleSize (:363-380),
 * HalfImageSize (:390-408), SubtractImage scalar (:460-464), fed like FEAT_SIFT_CPU::process (:86-90). */
static void sift_run(kplist *L, const uint8_t *gray, int height, int width, int double_size, int dbg_octave, float *dbg_gauss, float *dbg_dog) {
	const float peak_thresh = 0.04f / (float)SIFT_SCALES;
	simg org = img_new(height, width);
	for (int y = 0; y < height; ++y)
		for (int x = 0; x < width; ++x) org.px[(size_t)y * width + x] = (float)(((float)gray[(size_t)width * y + x]) * 1. / 255.);
	simg cur;
	float fscale = 1.0f;
	if (double_size) {
		int rows = height, cols = width, nr = 2 * rows - 2, nc = 2 * cols - 2;
		cur = img_new(nr, nc);
		for (int i = 0; i < rows - 1; ++i) {
			const float *ps = org.px + (size_t)i * cols;
			float *pd = cur.px + (size_t)(2 * i) * nc;
			for (int j = 0; j < cols - 1; ++j) {
				pd[2 * j] = ps[j];
				pd[nc + 2 * j] = 0.5f * (ps[j] + ps[cols + j]);
				pd[2 * j + 1] = 0.5f * (ps[j] + ps[j + 1]);
				pd[nc + 2 * j + 1] = 0.25f * (ps[j] + ps[j + 1] + ps[cols + j] + ps[cols + j + 1]);
			}
		}
		fscale = 0.5f;
	} else {
		cur = img_new(height, width);
		memcpy(cur.px, org.px, sizeof(float) * height * width);
	}
	free(org.px);
	float fnewscale = double_size ? 1.0f : 0.5f;
	if (SIFT_INIT_SIGMA > fnewscale) gaussian_blur(cur, cur, sqrtf(SIFT_INIT_SIGMA * SIFT_INIT_SIGMA - fnewscale * fnewscale));

	size_t cap = (size_t)cur.rows * cur.cols;
	simg gaus[SIFT_SCALES + 3], dog[SIFT_SCALES + 2], grad, orim;
	for (int i = 1; i < SIFT_SCALES + 3; ++i) gaus[i] = img_new(cur.rows, cur.cols);
	for (int i = 0; i < SIFT_SCALES + 2; ++i) dog[i] = img_new(cur.rows, cur.cols);
	grad = img_new(cur.rows, cur.cols); orim = img_new(cur.rows, cur.cols);
	char *maxmin = (char *)malloc(cap);
	float *first = cur.px;
	int octave = 0;
	float *prev_half = NULL;
	while (cur.rows > 12 && cur.cols > 12) {
		int rows = cur.rows, cols = cur.cols;
		float fwidth = powf(2.0f, 1.0f / (float)SIFT_SCALES);
		float fincsigma = sqrtf(fwidth * fwidth - 1.0f);
		gaus[0] = cur;
		float sigma = SIFT_INIT_SIGMA;
		for (int i = 1; i < SIFT_SCALES + 3; ++i) {
			gaus[i].rows = rows; gaus[i].cols = cols;
			gaussian_blur(gaus[i], gaus[i - 1], fincsigma * sigma);
			dog[i - 1].rows = rows; dog[i - 1].cols = cols;
			for (size_t k = 0; k < (size_t)rows * cols; ++k) dog[i - 1].px[k] = gaus[i - 1].px[k] - gaus[i].px[k];
			sigma *= fwidth;
		}
		if (octave == dbg_octave) {
			if (dbg_gauss) for (int i = 0; i < SIFT_SCALES + 3; ++i) memcpy(dbg_gauss + (size_t)i * rows * cols, gaus[i].px, sizeof(float) * rows * cols);
			if (dbg_dog) for (int i = 0; i < SIFT_SCALES + 2; ++i) memcpy(dbg_dog + (size_t)i * rows * cols, dog[i].px, sizeof(float) * rows * cols);
		}
		grad.rows = orim.rows = rows; grad.cols = orim.cols = cols;
		memset(maxmin, 0, (size_t)rows * cols);
		L->octave = octave;
		for (int index = 1; index < SIFT_SCALES + 1; ++index) {
			grad_ori(gaus[index], grad, orim);
			for (int r = 5; r < rows - 5; ++r) {
				const float *dp = dog[index].px + (size_t)r * cols;
				for (int c = 5; c < cols - 5; ++c) {
					float fval = dp[c];
					if (fabsf(fval) > peak_thresh * 0.8f) {
						if (local_max_min(fval, dog[index], r, c) && local_max_min(fval, dog[index - 1], r, c) &&
						    local_max_min(fval, dog[index + 1], r, c) && not_on_edge(dog[index], r, c))
							interp_keypoint(L, dog, index, r, c, grad, orim, maxmin, fscale, peak_thresh, r, c);
					}
				}
			}
		}
		/* HalfImageSize(s_imgaus[Scales]) */
		int nr = rows >> 1, nc = cols >> 1;
		simg half = img_new(nr, nc);
		for (int hr = 0; hr < nr; ++hr)
			for (int hc = 0; hc < nc; ++hc) half.px[(size_t)hr * nc + hc] = gaus[SIFT_SCALES].px[(size_t)(2 * hr) * cols + 2 * hc];
		if (prev_half) free(prev_half);
		prev_half = half.px;
		cur = half;
		fscale += fscale;
		octave++;
	}
	if (prev_half) free(prev_half);
	free(first);
	for (int i = 1; i < SIFT_SCALES + 3; ++i) free(gaus[i].px);
	for (int i = 0; i < SIFT_SCALES + 2; ++i) free(dog[i].px);
	free(grad.px); free(orim.px); free(maxmin);
}

static int sift_emit(kplist *L, int max_kp, float *xy, float *scale_ori, float *desc) {
	int n = L->n < max_kp ? L->n : max_kp;
	for (int i = 0; i < n; ++i) {          /* list order = reverse creation order */
		int s = L->n - 1 - i;
		if (xy) { xy[2 * i] = L->xy[2 * s]; xy[2 * i + 1] = L->xy[2 * s + 1]; }
		if (scale_ori) { scale_ori[2 * i] = L->so[2 * s]; scale_ori[2 * i + 1] = L->so[2 * s + 1]; }
		if (desc) memcpy(desc + (size_t)128 * i, L->desc + (size_t)128 * s, sizeof(float) * 128);
	}
	int total = L->n;
	free(L->xy); free(L->so); free(L->desc);
	return total;
}

int mo_sift(const uint8_t *gray, int height, int width, int double_size, int max_kp, float *xy, float *scale_ori, float *desc) {
	kplist L; memset(&L, 0, sizeof(L));
	sift_run(&L, gray, height, width, double_size, -1, NULL, NULL);
	return sift_emit(&L, max_kp, xy, scale_ori, desc);
}

int mo_sift_debug(const uint8_t *gray, int height, int width, int double_size, int dbg_octave, float *dbg_gauss, float *dbg_dog,
                  int max_trace, mo_sift_trace *trace, int *n_trace) {
	kplist L; memset(&L, 0, sizeof(L));
	L.trace = (mo_sift_trace *)malloc(sizeof(mo_sift_trace) * 1024); L.cap_trace = 1024;
	sift_run(&L, gray, height, width, double_size, dbg_octave, dbg_gauss, dbg_dog);
	int nt = L.n_trace < max_trace ? L.n_trace : max_trace;
	if (trace) memcpy(trace, L.trace, sizeof(mo_sift_trace) * nt);
	if (n_trace) *n_trace = L.n_trace;
	free(L.trace);
	return sift_emit(&L, 0, NULL, NULL, NULL);
}
