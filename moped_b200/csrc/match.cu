// match.cu — MATCH step on B200: device-resident model-descriptor database, fp16 tcgen05 distance
// contraction with a fused per-query top-k epilogue, exact fp32 re-rank in the reference's summation
// order, a per-query exactness certificate with an exhaustive exact fallback, and the ratio test.
//
// Replaces MATCH_ANN_CPU::Update/process (moped2/libmoped/src/match/MATCH_ANN_CPU.hpp:72-109,136-178)
// and the ANN kd-tree search behind it (libs.tgz!ann_1.1.1/kd_search.cpp:89-210). Parity target: the
// reference matcher in exact mode (Quality=0): same two nearest rows, same squared distances bit for bit.
//
// Kernels
//   k_pack_tiles      fp32 rows -> fp16 operand tile images, pre-swizzled (128B swizzle, K-major) so one
//                     32 KiB cp.async.bulk lands a ready-to-MMA tile in shared memory
//   k_match_coarse    warp-specialised tcgen05 kernel: TMA producer / MMA issuer / 8 epilogue warps
//   k_match_rerank    exact distances of the coarse candidates, top-2, certificate
//   k_match_exact*    exhaustive exact scan (fallback for uncertified queries, and MC_MATCH_EXACT)
//   k_match_finalize  ratio test
#include "common.cuh"

#include <math_constants.h>

// MC_COARSE_DBG (experiments only, never in the shipped library): 1 = the epilogue hands the accumulator stage back without
// reading it (MMA + TMA pace alone), 2 = the epilogue reads TMEM but skips the top-k scan (adds the TMEM-read pace)
#ifndef MC_COARSE_DBG
#define MC_COARSE_DBG 0
#endif

namespace mc {

// =============================================================================================
// small PTX wrappers (sm_100a)
// =============================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	uint32_t ok;
	do {
		asm volatile(
		    "{\n\t.reg .pred p;\n\t"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t}"
		    : "=r"(ok)
		    : "r"(bar), "r"(parity)
		    : "memory");
	} while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 1-D bulk copy global -> shared (TMA engine, UBLKCP), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
	    ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
	    : "memory");
}
// mbarrier arrive when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 consecutive fp32 columns: thread t gets row (lane base + t), r[j] = column (col base + j)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
	uint32_t *u = reinterpret_cast<uint32_t *>(r);
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
	      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
	      "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
	      "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
	    : "r"(taddr)
	    : "memory");
}

// order-preserving float <-> uint map, so that atomicMax works on scores of either sign.
// 0 (a cleared buffer) decodes to -inf.
__device__ __forceinline__ uint32_t f2o(float f) {
	uint32_t u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float o2f(uint32_t o) {
	if (o <= 0x007FFFFFu) return -CUDART_INF_F;
	return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}

// =============================================================================================
// operand tile images
// =============================================================================================
// Tile image of 128 rows x 128 fp16 (K-major): two K atoms of 64 elements; inside an atom row r is a
// 128-byte line at (r/8)*1024 + (r%8)*128 whose 16-byte chunk c sits at chunk position c ^ (r%8)
// (the 128B swizzle tcgen05 smem descriptors expect). One thread writes one 16-byte chunk.
__global__ void k_pack_tiles(const float *__restrict__ src, int64_t n_rows, int64_t n_tiles, __half *__restrict__ img,
                             float *__restrict__ norm2 /* nullable: per-row sum of squares */) {
	int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	int64_t total = n_tiles * 2048;
	if (gid >= total) return;
	int c = (int)(gid & 7);
	int r = (int)((gid >> 3) & 127);
	int a = (int)((gid >> 10) & 1);
	int64_t t = gid >> 11;
	int64_t row = t * kTileRows + r;
	uint4 out = make_uint4(0, 0, 0, 0);
	if (row < n_rows) {
		const float4 *p = reinterpret_cast<const float4 *>(src + row * kD + a * 64 + c * 8);
		float4 v0 = p[0], v1 = p[1];
		__half2 h0 = __floats2half2_rn(v0.x, v0.y), h1 = __floats2half2_rn(v0.z, v0.w);
		__half2 h2 = __floats2half2_rn(v1.x, v1.y), h3 = __floats2half2_rn(v1.z, v1.w);
		out.x = *reinterpret_cast<uint32_t *>(&h0); out.y = *reinterpret_cast<uint32_t *>(&h1);
		out.z = *reinterpret_cast<uint32_t *>(&h2); out.w = *reinterpret_cast<uint32_t *>(&h3);
	}
	size_t off = (size_t)t * kTileBytes + (size_t)a * 16384 + (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 + (size_t)((c ^ (r & 7)) * 16);
	*reinterpret_cast<uint4 *>(reinterpret_cast<char *>(img) + off) = out;
	if (norm2 && a == 0 && c == 0 && row < n_rows) {
		const float *p = src + row * kD;
		float s = 0.f;
		for (int d = 0; d < kD; d++) s = fmaf(p[d], p[d], s);
		norm2[row] = s;
	}
}

__global__ void k_minmax(const float *__restrict__ v, int64_t n, float *__restrict__ out2) {
	float lo = CUDART_INF_F, hi = -CUDART_INF_F;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		lo = fminf(lo, v[i]); hi = fmaxf(hi, v[i]);
	}
	for (int o = 16; o; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
	if ((threadIdx.x & 31) == 0) {
		atomicMin(reinterpret_cast<int *>(&out2[0]), __float_as_int(lo));   // norms are positive: int order == float order
		atomicMax(reinterpret_cast<int *>(&out2[1]), __float_as_int(hi));
	}
}

// =============================================================================================
// coarse pass: fp16 tcgen05 contraction with fused per-query top-k
// =============================================================================================
constexpr int kStages = 4;              // B-operand ring depth (4 x 32 KiB)
constexpr int kCoarseThreads = 384;     // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warp 3 idle, warps 4..11 epilogue
constexpr uint32_t kSmemA = 0;
constexpr uint32_t kSmemB = 2 * kTileBytes;
constexpr uint32_t kSmemBar = kSmemB + kStages * kTileBytes;
constexpr uint32_t kCoarseSmemBytes = kSmemBar + 256 + 1024;   // + barriers + alignment slack

// UMMA shared-memory descriptor: K-major, 128B swizzle, 8-row groups 1024 B apart (SBO), version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
	return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
	       ((uint64_t)2 << 61);
}
// instruction descriptor: D=f32, A=B=f16, both K-major, N=128, M=128
constexpr uint32_t kIdesc = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void topk_insert(float (&ts)[kTopK], int (&ti)[kTopK], float v, int row) {
	ts[kTopK - 1] = v; ti[kTopK - 1] = row;
#pragma unroll
	for (int j = kTopK - 1; j > 0; j--) {
		bool sw = ts[j] > ts[j - 1];
		float a = ts[j], b = ts[j - 1];
		int ia = ti[j], ib = ti[j - 1];
		ts[j - 1] = sw ? a : b; ts[j] = sw ? b : a;
		ti[j - 1] = sw ? ia : ib; ti[j] = sw ? ib : ia;
	}
}

// One chunk of 32 accumulator columns of this thread's query. Fast path: 4 group maxima against the
// running threshold; slow path (rare once the threshold has warmed up): insertion of the survivors.
__device__ __forceinline__ void scan_chunk(const float (&r)[32], int row_base, float &tau, float (&ts)[kTopK], int (&ti)[kTopK]) {
	float m0 = fmaxf(fmaxf(fmaxf(r[0], r[1]), fmaxf(r[2], r[3])), fmaxf(fmaxf(r[4], r[5]), fmaxf(r[6], r[7])));
	float m1 = fmaxf(fmaxf(fmaxf(r[8], r[9]), fmaxf(r[10], r[11])), fmaxf(fmaxf(r[12], r[13]), fmaxf(r[14], r[15])));
	float m2 = fmaxf(fmaxf(fmaxf(r[16], r[17]), fmaxf(r[18], r[19])), fmaxf(fmaxf(r[20], r[21]), fmaxf(r[22], r[23])));
	float m3 = fmaxf(fmaxf(fmaxf(r[24], r[25]), fmaxf(r[26], r[27])), fmaxf(fmaxf(r[28], r[29]), fmaxf(r[30], r[31])));
	float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
	if (m > tau) {
		// stage only the groups that hold a survivor, remember which elements beat the threshold, and
		// visit just those (typically one): the slow path must stay short, a warp pays it for all 32 queries
		float buf[32];
		unsigned mask = 0;
#define MC_STAGE_GROUP(g, mg)                                                   \
		if (mg > tau) {                                                         \
			_Pragma("unroll") for (int j = 0; j < 8; j++) {                     \
				buf[8 * g + j] = r[8 * g + j];                                  \
				mask |= (r[8 * g + j] > tau) ? (1u << (8 * g + j)) : 0u;        \
			}                                                                   \
		}
		MC_STAGE_GROUP(0, m0) MC_STAGE_GROUP(1, m1) MC_STAGE_GROUP(2, m2) MC_STAGE_GROUP(3, m3)
#undef MC_STAGE_GROUP
		while (mask) {
			const int j = __ffs(mask) - 1;
			mask &= mask - 1;
			const float v = buf[j];
			if (v > tau) {
				topk_insert(ts, ti, v, row_base + j);
				tau = fmaxf(tau, ts[kTopK - 1]);
			}
		}
	}
}

__device__ __forceinline__ void mask_tail(float (&r)[32], int64_t row_base, int64_t n_rows) {
#pragma unroll
	for (int j = 0; j < 32; j++)
		if (row_base + j >= n_rows) r[j] = -CUDART_INF_F;
}

// grid = (query tiles of 256, DB splits). Each CTA keeps its 256 queries' fp16 image resident in shared
// memory (A operand, 64 KiB) and streams its split of the DB tile images through a 4-stage ring (B operand).
// Per DB tile: 2 x 8 tcgen05.mma (M=128, N=128, K=16) into one of two TMEM accumulator stages
// (2 halves x 128 columns each); the 8 epilogue warps read the finished stage back with tcgen05.ld,
// one query row per thread, and keep that query's top-k (score, row) in registers.
__global__ void __launch_bounds__(kCoarseThreads, 1)
k_match_coarse(const __half *__restrict__ q_img, const __half *__restrict__ db_img, int64_t n_tiles, int64_t n_rows,
               int tiles_per_split, int n_splits, uint32_t *__restrict__ g_tau,
               float *__restrict__ cand_score, int32_t *__restrict__ cand_row) {
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int mtile = blockIdx.x, split = blockIdx.y;
	const int64_t t0 = (int64_t)split * tiles_per_split;
	const int64_t t1 = (t0 + tiles_per_split < n_tiles) ? t0 + tiles_per_split : n_tiles;
	const int ntiles = (int)(t1 > t0 ? t1 - t0 : 0);

	// barrier block: full[kStages], empty[kStages], a_full, tmem_full[2], tmem_empty[2], tmem base
	const uint32_t bar0 = smem_base + kSmemBar;
	auto bar_full = [&](int s) { return bar0 + 8u * s; };
	auto bar_empty = [&](int s) { return bar0 + 8u * (kStages + s); };
	const uint32_t bar_a = bar0 + 8u * (2 * kStages);
	auto bar_tfull = [&](int s) { return bar0 + 8u * (2 * kStages + 1 + s); };
	auto bar_tempty = [&](int s) { return bar0 + 8u * (2 * kStages + 3 + s); };
	volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem_gen + kSmemBar + 8 * (2 * kStages + 5));

	if (ntiles > 0) {
		if (threadIdx.x == 0) {
			for (int s = 0; s < kStages; s++) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
			mbar_init(bar_a, 1);
			for (int s = 0; s < 2; s++) { mbar_init(bar_tfull(s), 1); mbar_init(bar_tempty(s), 8); }
			fence_barrier_init();
		}
		if (warp == 2) tmem_alloc(smem_base + kSmemBar + 8 * (2 * kStages + 5), 512);
		tc_fence_before();
		__syncthreads();
		tc_fence_after();
	}
	const uint32_t tmem_base = ntiles > 0 ? *tmem_slot : 0;

	if (ntiles > 0 && warp == 0) {
		// ===== TMA producer =====
		if (lane == 0) {
			mbar_expect_tx(bar_a, 2 * kTileBytes);
			bulk_g2s(smem_base + kSmemA, reinterpret_cast<const char *>(q_img) + (size_t)mtile * 2 * kTileBytes, 2 * kTileBytes, bar_a);
			for (int i = 0; i < ntiles; i++) {
				int b = i % kStages;
				mbar_wait(bar_empty(b), ((i / kStages) & 1) ^ 1);
				mbar_expect_tx(bar_full(b), kTileBytes);
				bulk_g2s(smem_base + kSmemB + b * kTileBytes, reinterpret_cast<const char *>(db_img) + (size_t)(t0 + i) * kTileBytes,
				         kTileBytes, bar_full(b));
			}
		}
		__syncwarp();
	} else if (ntiles > 0 && warp == 1) {
		// ===== MMA issuer (one thread) =====
		if (lane == 0) {
			mbar_wait(bar_a, 0);
			for (int i = 0; i < ntiles; i++) {
				int s = i & 1, b = i % kStages;
				mbar_wait(bar_tempty(s), ((i >> 1) & 1) ^ 1);
				mbar_wait(bar_full(b), (i / kStages) & 1);
				tc_fence_after();
#pragma unroll
				for (int h = 0; h < 2; h++) {
					const uint32_t d_tmem = tmem_base + (uint32_t)(s * 256 + h * 128);
#pragma unroll
					for (int k = 0; k < 8; k++) {
						const uint32_t koff = (uint32_t)((k >> 2) * 16384 + (k & 3) * 32);
						uint64_t a_desc = make_desc(smem_base + kSmemA + h * kTileBytes + koff);
						uint64_t b_desc = make_desc(smem_base + kSmemB + b * kTileBytes + koff);
						umma_f16(d_tmem, a_desc, b_desc, kIdesc, k > 0 ? 1u : 0u);
					}
				}
				umma_commit(bar_empty(b));     // smem stage reusable once these MMAs have read it
				umma_commit(bar_tfull(s));     // accumulator stage complete
			}
		}
		__syncwarp();
	} else if (warp >= 4) {
		// ===== epilogue: one query per thread =====
		const int e = warp - 4;
		const int quarter = e & 3, half = e >> 2;          // TMEM lane quarter == warp % 4
		const int qid = mtile * kMTile + half * 128 + quarter * 32 + lane;
		float ts[kTopK]; int ti[kTopK];
#pragma unroll
		for (int j = 0; j < kTopK; j++) { ts[j] = -CUDART_INF_F; ti[j] = -1; }
		float tau = -CUDART_INF_F;
		float published = -CUDART_INF_F;
		float ra[32], rb[32];
		uint32_t og = 0;                                   // other CTAs' threshold for this query, fetched one tile ahead
		for (int i = 0; i < ntiles; i++) {
			const int s = i & 1;
			tau = fmaxf(tau, o2f(og));
			og = __ldcg(&g_tau[qid]);                      // consumed at the top of the next tile: its L2 latency stays hidden
			mbar_wait(bar_tfull(s), (i >> 1) & 1);
			tc_fence_after();
			const int64_t row0 = (t0 + i) * kTileRows;
			const bool tail = row0 + kTileRows > n_rows;
			const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * 256 + half * 128);
#if MC_COARSE_DBG == 1
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_tempty(s));
			(void)taddr; (void)tail; (void)ra; (void)rb;
#elif MC_COARSE_DBG == 2
			tmem_ld32(taddr, ra); tmem_ld_wait();
			tmem_ld32(taddr + 32, rb); tmem_ld_wait();
			tau = fmaxf(tau, ra[0] + rb[31]);
			tmem_ld32(taddr + 64, ra); tmem_ld_wait();
			tmem_ld32(taddr + 96, rb); tmem_ld_wait();
			tau = fmaxf(tau, ra[0] + rb[31]);
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_tempty(s));
			(void)tail;
#else
			tmem_ld32(taddr, ra);
			tmem_ld_wait();
			tmem_ld32(taddr + 32, rb);
			if (tail) mask_tail(ra, row0, n_rows);
			scan_chunk(ra, (int)row0, tau, ts, ti);
			tmem_ld_wait();
			tmem_ld32(taddr + 64, ra);
			if (tail) mask_tail(rb, row0 + 32, n_rows);
			scan_chunk(rb, (int)row0 + 32, tau, ts, ti);
			tmem_ld_wait();
			tmem_ld32(taddr + 96, rb);
			if (tail) mask_tail(ra, row0 + 64, n_rows);
			scan_chunk(ra, (int)row0 + 64, tau, ts, ti);
			tmem_ld_wait();
			// all TMEM reads of this stage are done: hand it back to the MMA warp
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_tempty(s));
			if (tail) mask_tail(rb, row0 + 96, n_rows);
			scan_chunk(rb, (int)row0 + 96, tau, ts, ti);
#endif
			if (ts[kTopK - 1] > published) {               // publish this CTA's k-th best (a valid global lower bound)
				published = ts[kTopK - 1];
				atomicMax(&g_tau[qid], f2o(published));
			}
		}
		const size_t o = ((size_t)qid * n_splits + split) * kTopK;
#pragma unroll
		for (int j = 0; j < kTopK; j++) { cand_score[o + j] = ts[j]; cand_row[o + j] = ti[j]; }
	}

	if (ntiles > 0) {
		tc_fence_before();
		__syncthreads();
		if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
	}
}

// =============================================================================================
// exact arithmetic of the reference: dist = sum_d (q[d]-p[d])^2, sequential, product and sum each
// rounded to fp32 (VMULSS + VADDSS in the compiled reference, libs.tgz!ann_1.1.1/kd_search.cpp:188-199)
// =============================================================================================
struct Top2 {
	float d0, d1;
	int32_t i0, i1;
};
__device__ __forceinline__ bool closer(float d, int32_t i, float e, int32_t j) { return d < e || (d == e && (uint32_t)i < (uint32_t)j); }
__device__ __forceinline__ void top2_push(Top2 &t, float d, int32_t i) {
	if (i < 0) return;
	if (closer(d, i, t.d0, t.i0)) { t.d1 = t.d0; t.i1 = t.i0; t.d0 = d; t.i0 = i; }
	else if (closer(d, i, t.d1, t.i1)) { t.d1 = d; t.i1 = i; }
}
__device__ __forceinline__ void top2_warp_reduce(Top2 &t) {
#pragma unroll
	for (int o = 16; o; o >>= 1) {
		float d0 = __shfl_xor_sync(0xffffffffu, t.d0, o), d1 = __shfl_xor_sync(0xffffffffu, t.d1, o);
		int32_t i0 = __shfl_xor_sync(0xffffffffu, t.i0, o), i1 = __shfl_xor_sync(0xffffffffu, t.i1, o);
		top2_push(t, d0, i0);
		top2_push(t, d1, i1);
	}
}
__device__ __forceinline__ Top2 top2_empty() {
	Top2 t; t.d0 = t.d1 = 3.402823466e+38f /* PQ_NULL_KEY = ANN_DIST_INF: a missing neighbour */; t.i0 = t.i1 = -1; return t;
}

// Exact re-rank of the coarse candidates; one warp per query.
// Certificate: every row that is NOT a candidate scored <= T = final shared threshold of this query in the
// fp16 pass, so its true dot product is <= T + E and its true squared distance is
// >= |q|^2 + min|p|^2 - 2 (T + E); if that bound exceeds the exact 2nd-best candidate distance the two
// nearest rows are proven to be among the candidates. E bounds the fp16 input rounding (2^-10 |q||p|),
// fp16 subnormals and the fp32 accumulation of the tensor core. Uncertified queries go to the exact scan.
__global__ void k_match_rerank(const float *__restrict__ q, const float *__restrict__ q_norm2, int Q, const float *__restrict__ db,
                               const float *__restrict__ cand_score, const int32_t *__restrict__ cand_row, int n_cand,
                               const uint32_t *__restrict__ g_tau, float db_n2_min, float db_n2_max, int64_t row_base,
                               int32_t *__restrict__ nn_row, float *__restrict__ nn_dist, int32_t *__restrict__ flag_list,
                               int32_t *__restrict__ flag_count, unsigned long long *__restrict__ nn_key) {
	__shared__ float qs[8][kD];
	const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int qi = blockIdx.x * 8 + w;
	if (qi >= Q) return;
	for (int d = lane; d < kD; d += 32) qs[w][d] = q[(size_t)qi * kD + d];
	__syncwarp();
	Top2 t = top2_empty();
	for (int c = lane; c < n_cand; c += 32) {
		int32_t row = cand_row[(size_t)qi * n_cand + c];
		if (row < 0) continue;
		const float4 *p = reinterpret_cast<const float4 *>(db + (size_t)row * kD);
		float dist = 0.f;
#pragma unroll 4
		for (int d4 = 0; d4 < kD / 4; d4++) {
			float4 v = __ldg(&p[d4]);
			float t0 = __fsub_rn(qs[w][4 * d4 + 0], v.x); dist = __fadd_rn(dist, __fmul_rn(t0, t0));
			float t1 = __fsub_rn(qs[w][4 * d4 + 1], v.y); dist = __fadd_rn(dist, __fmul_rn(t1, t1));
			float t2 = __fsub_rn(qs[w][4 * d4 + 2], v.z); dist = __fadd_rn(dist, __fmul_rn(t2, t2));
			float t3 = __fsub_rn(qs[w][4 * d4 + 3], v.w); dist = __fadd_rn(dist, __fmul_rn(t3, t3));
		}
		top2_push(t, dist, row);
	}
	top2_warp_reduce(t);
	if (lane == 0) {
		const float T = o2f(g_tau[qi]);
		const float qn2 = q_norm2[qi];
		const float E = 1.1e-3f * sqrtf(qn2 * db_n2_max) + 1e-5f;
		const float bound = qn2 + db_n2_min - 2.f * (T + E) - 1e-4f;
		const bool certified = (t.i1 >= 0) && (T == -CUDART_INF_F || t.d1 < bound);
		nn_row[2 * qi] = t.i0 >= 0 ? (int32_t)(t.i0 + row_base) : -1;
		nn_row[2 * qi + 1] = t.i1 >= 0 ? (int32_t)(t.i1 + row_base) : -1;
		nn_dist[2 * qi] = t.d0; nn_dist[2 * qi + 1] = t.d1;
		if (!certified) {
			flag_list[atomicAdd(flag_count, 1)] = qi;
			nn_key[2 * (size_t)qi] = ~0ull; nn_key[2 * (size_t)qi + 1] = ~0ull;
		}
	}
}

// Two-smallest accumulator over 64-bit keys (distance bits << 32 | row): slot 0 takes every key by
// atomicMin; whichever of (old slot 0, key) loses goes to slot 1 by atomicMin. Keys are unique, the
// global minimum is never a loser and the global second minimum always is, so slot 1 ends as the 2nd.
__device__ __forceinline__ void key2_push(unsigned long long *slot, float d, int32_t row) {
	if (row < 0) return;
	unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (uint32_t)row;
	unsigned long long old = atomicMin(&slot[0], key);
	unsigned long long loser = old < key ? key : old;
	if (loser != ~0ull) atomicMin(&slot[1], loser);
}

// Exhaustive exact scan. Work item = (group of 8 listed queries, row chunk); persistent grid.
// Each thread walks rows (stride blockDim) of the chunk and carries the 8 queries' running sums through
// the descriptor in order, 16 dimensions at a time; a block-reduced top-2 per query is merged into the
// query's global two-smallest slots.
constexpr int kExactQ = 8;
constexpr int kExactThreads = 256;
constexpr int kMaxAnyD = 4096;          // longest descriptor the generic exact scan stages in shared memory (8 x 4096 floats = 128 KiB)
__global__ void __launch_bounds__(kExactThreads)
k_match_exact(const float *__restrict__ q, const int32_t *__restrict__ list, const int32_t *__restrict__ list_count, int list_all,
              const float *__restrict__ db, int64_t n_rows, int n_chunks, int chunk_rows,
              unsigned long long *__restrict__ nn_key) {
	__shared__ __align__(16) float qs[kExactQ][kD];
	__shared__ Top2 red[kExactQ][kExactThreads / 32];
	const int n_list = list ? *list_count : list_all;
	const int n_groups = (n_list + kExactQ - 1) / kExactQ;
	const int64_t n_items = (int64_t)n_groups * n_chunks;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
		const int g = (int)(item / n_chunks), ch = (int)(item % n_chunks);
		__syncthreads();
		for (int x = threadIdx.x; x < kExactQ * kD; x += blockDim.x) {
			int k = x / kD, li = g * kExactQ + k;
			int qi = li < n_list ? (list ? list[li] : li) : -1;
			qs[k][x % kD] = qi >= 0 ? q[(size_t)qi * kD + (x % kD)] : 0.f;
		}
		__syncthreads();
		Top2 t[kExactQ];
#pragma unroll
		for (int k = 0; k < kExactQ; k++) t[k] = top2_empty();
		const int64_t r_lo = (int64_t)ch * chunk_rows;
		const int64_t r_hi = r_lo + chunk_rows < n_rows ? r_lo + chunk_rows : n_rows;
		for (int64_t row = r_lo + threadIdx.x; row < r_hi; row += blockDim.x) {
			const float4 *p = reinterpret_cast<const float4 *>(db + (size_t)row * kD);
			float dist[kExactQ];
#pragma unroll
			for (int k = 0; k < kExactQ; k++) dist[k] = 0.f;
#pragma unroll 2
			for (int b = 0; b < kD / 16; b++) {
				float4 v0 = __ldg(&p[4 * b]), v1 = __ldg(&p[4 * b + 1]), v2 = __ldg(&p[4 * b + 2]), v3 = __ldg(&p[4 * b + 3]);
				const float pv[16] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w };
#pragma unroll
				for (int k = 0; k < kExactQ; k++) {
					const float4 *qq = reinterpret_cast<const float4 *>(&qs[k][16 * b]);
					float4 a0 = qq[0], a1 = qq[1], a2 = qq[2], a3 = qq[3];
					const float qv[16] = { a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w };
					float d = dist[k];
#pragma unroll
					for (int j = 0; j < 16; j++) { float tt = __fsub_rn(qv[j], pv[j]); d = __fadd_rn(d, __fmul_rn(tt, tt)); }
					dist[k] = d;
				}
			}
#pragma unroll
			for (int k = 0; k < kExactQ; k++) top2_push(t[k], dist[k], (int32_t)row);
		}
#pragma unroll
		for (int k = 0; k < kExactQ; k++) {
			top2_warp_reduce(t[k]);
			if (lane == 0) red[k][warp] = t[k];
		}
		__syncthreads();
		if (threadIdx.x < kExactQ) {
			const int k = threadIdx.x, li = g * kExactQ + k;
			if (li < n_list) {
				Top2 r = red[k][0];
				for (int ww = 1; ww < kExactThreads / 32; ww++) { top2_push(r, red[k][ww].d0, red[k][ww].i0); top2_push(r, red[k][ww].d1, red[k][ww].i1); }
				const int qi = list ? list[li] : li;
				key2_push(nn_key + 2 * (size_t)qi, r.d0, r.i0);
				key2_push(nn_key + 2 * (size_t)qi, r.d1, r.i1);
			}
		}
	}
}

// The same exhaustive scan for ANY descriptor length (MATCH_ANN_CPU's constructor takes DescriptorSize, e.g. SURF-64,
// moped2/libmoped/src/match/MATCH_ANN_CPU.hpp:113): runtime D, the group's queries staged in dynamic shared memory
// (kExactQ x D floats), every distance summed over d = 0..D-1 in order with unfused multiply and add like the 128-d kernel.
__global__ void __launch_bounds__(kExactThreads)
k_match_exact_any(const float *__restrict__ q, int D, const int32_t *__restrict__ list, const int32_t *__restrict__ list_count, int list_all,
                  const float *__restrict__ db, int64_t n_rows, int n_chunks, int chunk_rows, unsigned long long *__restrict__ nn_key) {
	extern __shared__ __align__(16) float qs_any[];            // [kExactQ][D]
	__shared__ Top2 red[kExactQ][kExactThreads / 32];
	const int n_list = list ? *list_count : list_all;
	const int n_groups = (n_list + kExactQ - 1) / kExactQ;
	const int64_t n_items = (int64_t)n_groups * n_chunks;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const bool vec4 = (D & 3) == 0;
	for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
		const int g = (int)(item / n_chunks), ch = (int)(item % n_chunks);
		__syncthreads();
		for (int x = threadIdx.x; x < kExactQ * D; x += blockDim.x) {
			const int k = x / D, d = x - k * D, li = g * kExactQ + k;
			const int qi = li < n_list ? (list ? list[li] : li) : -1;
			qs_any[x] = qi >= 0 ? q[(size_t)qi * D + d] : 0.f;
		}
		__syncthreads();
		Top2 t[kExactQ];
#pragma unroll
		for (int k = 0; k < kExactQ; k++) t[k] = top2_empty();
		const int64_t r_lo = (int64_t)ch * chunk_rows;
		const int64_t r_hi = r_lo + chunk_rows < n_rows ? r_lo + chunk_rows : n_rows;
		for (int64_t row = r_lo + threadIdx.x; row < r_hi; row += blockDim.x) {
			const float *p = db + (size_t)row * D;
			float dist[kExactQ];
#pragma unroll
			for (int k = 0; k < kExactQ; k++) dist[k] = 0.f;
			if (vec4) {
				for (int d = 0; d < D; d += 4) {
					const float4 v = __ldg(reinterpret_cast<const float4 *>(p + d));
#pragma unroll
					for (int k = 0; k < kExactQ; k++) {
						const float4 a = *reinterpret_cast<const float4 *>(&qs_any[k * D + d]);
						float dd = dist[k], tt;
						tt = __fsub_rn(a.x, v.x); dd = __fadd_rn(dd, __fmul_rn(tt, tt));
						tt = __fsub_rn(a.y, v.y); dd = __fadd_rn(dd, __fmul_rn(tt, tt));
						tt = __fsub_rn(a.z, v.z); dd = __fadd_rn(dd, __fmul_rn(tt, tt));
						tt = __fsub_rn(a.w, v.w); dd = __fadd_rn(dd, __fmul_rn(tt, tt));
						dist[k] = dd;
					}
				}
			} else {
				for (int d = 0; d < D; d++) {
					const float v = __ldg(p + d);
#pragma unroll
					for (int k = 0; k < kExactQ; k++) { const float tt = __fsub_rn(qs_any[k * D + d], v); dist[k] = __fadd_rn(dist[k], __fmul_rn(tt, tt)); }
				}
			}
#pragma unroll
			for (int k = 0; k < kExactQ; k++) top2_push(t[k], dist[k], (int32_t)row);
		}
#pragma unroll
		for (int k = 0; k < kExactQ; k++) {
			top2_warp_reduce(t[k]);
			if (lane == 0) red[k][warp] = t[k];
		}
		__syncthreads();
		if (threadIdx.x < kExactQ) {
			const int k = threadIdx.x, li = g * kExactQ + k;
			if (li < n_list) {
				Top2 r = red[k][0];
				for (int ww = 1; ww < kExactThreads / 32; ww++) { top2_push(r, red[k][ww].d0, red[k][ww].i0); top2_push(r, red[k][ww].d1, red[k][ww].i1); }
				const int qi = list ? list[li] : li;
				key2_push(nn_key + 2 * (size_t)qi, r.d0, r.i0);
				key2_push(nn_key + 2 * (size_t)qi, r.d1, r.i1);
			}
		}
	}
}

// decode the two-smallest slots of every listed query into (global row id, distance)
__global__ void k_match_exact_decode(const int32_t *__restrict__ list, const int32_t *__restrict__ list_count, int list_all,
                                     const unsigned long long *__restrict__ nn_key, int64_t row_base,
                                     int32_t *__restrict__ nn_row, float *__restrict__ nn_dist) {
	const int n_list = list ? *list_count : list_all;
	for (int li = blockIdx.x * blockDim.x + threadIdx.x; li < n_list; li += gridDim.x * blockDim.x) {
		const int qi = list ? list[li] : li;
#pragma unroll
		for (int j = 0; j < 2; j++) {
			unsigned long long k = nn_key[2 * (size_t)qi + j];
			bool ok = k != ~0ull;
			nn_row[2 * qi + j] = ok ? (int32_t)((int64_t)(uint32_t)(k & 0xffffffffull) + row_base) : -1;
			nn_dist[2 * qi + j] = ok ? __uint_as_float((uint32_t)(k >> 32)) : 3.402823466e+38f;
		}
	}
}

// ratio test on squared distances, fp32 division like `ds[0]/ds[1] < Ratio` (MATCH_ANN_CPU.hpp:165)
__global__ void k_match_finalize(const int32_t *__restrict__ nn_row, const float *__restrict__ nn_dist, int Q, float ratio,
                                 uint8_t *__restrict__ accepted) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= Q) return;
	bool ok = nn_row[2 * i] >= 0 && nn_row[2 * i + 1] >= 0 && __fdiv_rn(nn_dist[2 * i], nn_dist[2 * i + 1]) < ratio;
	accepted[i] = ok ? 1 : 0;
}

// global top-2 over object shards: (smaller distance, then smaller global row id)
// shard s's rows / distances start `stride` elements after shard s-1's (2 Q for two separate arrays, 4 Q for packed blocks)
__global__ void k_match_merge(const int32_t *__restrict__ rows_all, const float *__restrict__ dist_all, size_t stride, int n_shards, int Q, float ratio,
                              int32_t *__restrict__ nn_row, float *__restrict__ nn_dist, uint8_t *__restrict__ accepted) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= Q) return;
	Top2 t = top2_empty();
	for (int s = 0; s < n_shards; s++) {
		size_t o = (size_t)s * stride + (size_t)i * 2;
		top2_push(t, dist_all[o], rows_all[o]);
		top2_push(t, dist_all[o + 1], rows_all[o + 1]);
	}
	nn_row[2 * i] = t.i0; nn_row[2 * i + 1] = t.i1;
	nn_dist[2 * i] = t.d0; nn_dist[2 * i + 1] = t.d1;
	accepted[i] = (t.i0 >= 0 && t.i1 >= 0 && __fdiv_rn(t.d0, t.d1) < ratio) ? 1 : 0;
}

__global__ void k_row_norm2(const float *__restrict__ src, int n, float *__restrict__ norm2) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float *p = src + (size_t)i * kD;
	float s = 0.f;
	for (int d = 0; d < kD; d++) s = fmaf(p[d], p[d], s);
	norm2[i] = s;
}

// =============================================================================================
// host side
// =============================================================================================
// Function attributes are PER DEVICE: called by mc_create for the context's device (a process may hold contexts on
// several GPUs), never behind a process-wide flag.
mc_status match_configure_device(mc_ctx *ctx) {
	MC_CUDA(cudaFuncSetAttribute(k_match_coarse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoarseSmemBytes));
	MC_CUDA(cudaFuncSetAttribute(k_match_exact_any, cudaFuncAttributeMaxDynamicSharedMemorySize, kExactQ * kMaxAnyD * (int)sizeof(float)));
	return MC_OK;
}

mc_status db_build_images(mc_ctx *ctx) {
	if (ctx->D != kD) return MC_OK;     // tensor path needs D == 128; other lengths use the exact scan
	ctx->n_tiles = (ctx->n_rows + kTileRows - 1) / kTileRows;
	MC_CUDA(cudaMalloc(&ctx->d_db_img, (size_t)ctx->n_tiles * kTileBytes));
	float *d_norm2 = nullptr, *d_mm = nullptr;
	MC_CUDA(cudaMalloc(&d_norm2, sizeof(float) * (size_t)(ctx->n_rows + 1)));
	MC_CUDA(cudaMalloc(&d_mm, 2 * sizeof(float)));
	const float init[2] = { 3.0e38f, 0.f };
	MC_CUDA(cudaMemcpyAsync(d_mm, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
	int64_t total = ctx->n_tiles * 2048;
	k_pack_tiles<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_db, ctx->n_rows, ctx->n_tiles, ctx->d_db_img, d_norm2);
	MC_LAUNCH_CHECK();
	k_minmax<<<296, 256, 0, ctx->stream>>>(d_norm2, ctx->n_rows, d_mm);
	MC_LAUNCH_CHECK();
	float mm[2];
	MC_CUDA(cudaMemcpyAsync(mm, d_mm, sizeof mm, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->db_norm2_min = mm[0]; ctx->db_norm2_max = mm[1];
	MC_CUDA(cudaFree(d_norm2));
	MC_CUDA(cudaFree(d_mm));
	return MC_OK;
}

static mc_status exact_scan(mc_ctx *ctx, const float *d_q, int Q, const int32_t *d_list, const int32_t *d_count, int32_t *d_nn_row, float *d_nn_dist) {
	// chunks: enough items to fill the machine even for a single group of queries
	int n_chunks = ctx->num_sms * 2;
	int chunk_rows = (int)((ctx->n_rows + n_chunks - 1) / n_chunks);
	if (chunk_rows < kExactThreads) chunk_rows = kExactThreads;
	n_chunks = (int)((ctx->n_rows + chunk_rows - 1) / chunk_rows);
	if (!d_list) MC_CUDA(cudaMemsetAsync(ctx->nn_key.p, 0xFF, sizeof(unsigned long long) * 2 * (size_t)Q, ctx->stream));
	if (ctx->D == kD)
		k_match_exact<<<ctx->num_sms * 4, kExactThreads, 0, ctx->stream>>>(d_q, d_list, d_count, Q, ctx->d_db, ctx->n_rows, n_chunks, chunk_rows,
		                                                                 (unsigned long long *)ctx->nn_key.p);
	else
		k_match_exact_any<<<ctx->num_sms * 4, kExactThreads, sizeof(float) * kExactQ * (size_t)ctx->D, ctx->stream>>>(
		    d_q, ctx->D, d_list, d_count, Q, ctx->d_db, ctx->n_rows, n_chunks, chunk_rows, (unsigned long long *)ctx->nn_key.p);
	MC_LAUNCH_CHECK();
	k_match_exact_decode<<<ctx->num_sms, 256, 0, ctx->stream>>>(d_list, d_count, Q, (const unsigned long long *)ctx->nn_key.p, ctx->row_base,
	                                                          d_nn_row, d_nn_dist);
	MC_LAUNCH_CHECK();
	return MC_OK;
}

mc_status match_device(mc_ctx *ctx, const float *d_q, int Q, float ratio, int mode, int32_t *d_nn_row, float *d_nn_dist, uint8_t *d_accepted) {
	if (!ctx->d_db) { ctx->err = "mc_match: no database uploaded"; return MC_ERR_STATE; }
	if (Q <= 0) return MC_OK;
	MC_TRY(reserve(ctx, ctx->nn_key, sizeof(unsigned long long) * 2 * (size_t)Q));
	// the tensor-core path is built for 128-d descriptors (SIFT); any other DescriptorSize takes the exhaustive exact scan —
	// the same result by definition (MC_MATCH_TENSOR only ever promises the exact scan's bits)
	ctx->last_match_q = Q;
	ctx->last_match_tensor = !(mode == MC_MATCH_EXACT || ctx->D != kD);
	if (mode == MC_MATCH_EXACT || ctx->D != kD) {
		MC_TRY(exact_scan(ctx, d_q, Q, nullptr, nullptr, d_nn_row, d_nn_dist));
	} else {
		const int n_mtiles = (Q + kMTile - 1) / kMTile;
		const int q_pad = n_mtiles * kMTile;
		// DB splits. Work item = (query tile, DB split), all items equally long. Few query tiles (one frame): a
		// single wave, floor(SMs / tiles) splits — a 149th CTA would run alone in a second wave. Many query tiles
		// (frame batches): several waves; keep at least 8 splits so that every query has >= 32 coarse candidates
		// (with fewer the exactness certificate starts to fail and queries fall to the exhaustive scan), and among
		// 8..64 splits take the count that wastes the least of the last wave.
		int n_splits = ctx->num_sms / n_mtiles;
		if (n_splits < 8) {
			double best = -1.0;
			for (int sp = 8; sp <= kMaxSplits; sp++) {
				const int64_t items = (int64_t)n_mtiles * sp;
				const int64_t waves = (items + ctx->num_sms - 1) / ctx->num_sms;
				const double eff = (double)items / (double)(waves * ctx->num_sms);
				if (eff > best + 0.01) { best = eff; n_splits = sp; }
			}
		}
		if (n_splits < 1) n_splits = 1;
		if (n_splits > kMaxSplits) n_splits = kMaxSplits;
		if ((int64_t)n_splits > ctx->n_tiles) n_splits = (int)ctx->n_tiles;
		int tiles_per_split = (int)((ctx->n_tiles + n_splits - 1) / n_splits);
		n_splits = (int)((ctx->n_tiles + tiles_per_split - 1) / tiles_per_split);
		const int n_cand = n_splits * kTopK;
		MC_TRY(reserve(ctx, ctx->q_img, (size_t)n_mtiles * 2 * kTileBytes));
		MC_TRY(reserve(ctx, ctx->q_norm2, sizeof(float) * q_pad));
		MC_TRY(reserve(ctx, ctx->tau, sizeof(uint32_t) * q_pad));
		MC_TRY(reserve(ctx, ctx->cand_score, sizeof(float) * (size_t)q_pad * n_cand));
		MC_TRY(reserve(ctx, ctx->cand_row, sizeof(int32_t) * (size_t)q_pad * n_cand));
		MC_TRY(reserve(ctx, ctx->flag_list, sizeof(int32_t) * q_pad));
		MC_TRY(reserve(ctx, ctx->flag_count, 256));
		MC_CUDA(cudaMemsetAsync(ctx->tau.p, 0, sizeof(uint32_t) * q_pad, ctx->stream));
		MC_CUDA(cudaMemsetAsync(ctx->flag_count.p, 0, sizeof(int32_t), ctx->stream));
		int64_t total = (int64_t)n_mtiles * 2 * 2048;
		k_pack_tiles<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_q, Q, (int64_t)n_mtiles * 2, (__half *)ctx->q_img.p, (float *)ctx->q_norm2.p);
		MC_LAUNCH_CHECK();
		dim3 grid(n_mtiles, n_splits);
		if (ctx->profile) MC_CUDA(cudaEventRecord(ctx->ev_coarse[0], ctx->stream));
		k_match_coarse<<<grid, kCoarseThreads, kCoarseSmemBytes, ctx->stream>>>((const __half *)ctx->q_img.p, ctx->d_db_img, ctx->n_tiles, ctx->n_rows,
		                                                                       tiles_per_split, n_splits, (uint32_t *)ctx->tau.p,
		                                                                       (float *)ctx->cand_score.p, (int32_t *)ctx->cand_row.p);
		MC_LAUNCH_CHECK();
		if (ctx->profile) { MC_CUDA(cudaEventRecord(ctx->ev_coarse[1], ctx->stream)); ctx->ev_valid = true; }
		k_match_rerank<<<(Q + 7) / 8, 256, 0, ctx->stream>>>(d_q, (const float *)ctx->q_norm2.p, Q, ctx->d_db, (const float *)ctx->cand_score.p,
		                                                    (const int32_t *)ctx->cand_row.p, n_cand, (const uint32_t *)ctx->tau.p,
		                                                    ctx->db_norm2_min, ctx->db_norm2_max, ctx->row_base, d_nn_row, d_nn_dist,
		                                                    (int32_t *)ctx->flag_list.p, (int32_t *)ctx->flag_count.p, (unsigned long long *)ctx->nn_key.p);
		MC_LAUNCH_CHECK();
		MC_TRY(exact_scan(ctx, d_q, Q, (const int32_t *)ctx->flag_list.p, (const int32_t *)ctx->flag_count.p, d_nn_row, d_nn_dist));
		ctx->last_stats[2] = n_cand; ctx->last_stats[3] = n_splits;
	}
	k_match_finalize<<<(Q + 255) / 256, 256, 0, ctx->stream>>>(d_nn_row, d_nn_dist, Q, ratio, d_accepted);
	MC_LAUNCH_CHECK();
	return MC_OK;
}

mc_status match_merge_device(mc_ctx *ctx, const int32_t *rows_all, const float *dist_all, size_t stride, int n_shards, int Q, float ratio,
                             int32_t *d_nn_row, float *d_nn_dist, uint8_t *d_accepted) {
	if (Q <= 0) return MC_OK;
	k_match_merge<<<(Q + 255) / 256, 256, 0, ctx->stream>>>(rows_all, dist_all, stride, n_shards, Q, ratio, d_nn_row, d_nn_dist, d_accepted);
	MC_LAUNCH_CHECK();
	return MC_OK;
}

} // namespace mc
