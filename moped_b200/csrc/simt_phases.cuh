// simt_phases.cuh — cooperative device code written as a sequence of PHASES, so that the same source also compiles for the host.
//
// A phase is `team.each(f)`: f(l) for every lane / thread l of the team, all of them finished (and their writes visible) before
// anything that follows. On the device that is  sync; f(my_index); sync  with __syncwarp (WarpTeam) or __syncthreads (BlockTeam);
// the leading sync also orders the reads of the code before the phase against the phase's writes. On the HOST it is a loop over
// the members — which is how tests/cpp/*_host.cpp run lm_exact.cuh, depth_pose.cuh and linkage_cached.cuh under g++ and compare
// them with the oracle on a machine without a GPU, visiting the members in ascending and in descending order (phx::g_host_order):
// a phase whose members depended on each other's order would show. Rules for code between phases: it runs on every member with
// identical values (control flow stays uniform) and reads, never writes, the state the phases share.
// Nothing here is a CPU path of the product: the library only instantiates these templates inside kernels.
#pragma once

#if defined(__CUDACC__)
#define PHX_FN __device__ __forceinline__
#define PHX_MEM __device__ __forceinline__
#else
#define PHX_FN static inline
#define PHX_MEM inline
#endif

namespace phx {

#if !defined(__CUDA_ARCH__)
static int g_host_order = 0;             // host emulation only: 0 = members ascending, 1 = descending
#endif

template <int W>
struct WarpTeam {                        // W consecutive lanes of one warp: 1, 8 (four teams per warp) or 32
	int lane;                            // index inside the team, 0..W-1
	int base;                            // the team's first lane inside its warp
	unsigned mask;                       // the team's lanes inside its warp (device only)
	PHX_MEM void init(int lane_in_warp) {
		lane = lane_in_warp % W;
		base = lane_in_warp - lane;
		mask = W >= 32 ? 0xffffffffu : ((1u << W) - 1u) << base;
	}
	// the value member 0 holds, on every member (keeps data-dependent control flow uniform inside the team)
	PHX_MEM int from_first(int v) const {
#if defined(__CUDA_ARCH__)
		return W > 1 ? __shfl_sync(mask, v, base) : v;
#else
		return v;
#endif
	}
	PHX_MEM void sync() const {
#if defined(__CUDA_ARCH__)
		if (W > 1) __syncwarp(mask);
#endif
	}
	template <class F> PHX_MEM void each(F f) const {
#if defined(__CUDA_ARCH__)
		sync();
		f(lane);
		sync();
#else
		if (g_host_order == 0) for (int l = 0; l < W; l++) f(l);
		else for (int l = W - 1; l >= 0; l--) f(l);
#endif
	}
};

template <int W>
struct BlockTeam {                       // the W threads of a CTA
	int tid;
	template <class F> PHX_MEM void each(F f) const {
#if defined(__CUDA_ARCH__)
		__syncthreads();
		f(tid);
		__syncthreads();
#else
		if (g_host_order == 0) for (int t = 0; t < W; t++) f(t);
		else for (int t = W - 1; t >= 0; t--) f(t);
#endif
	}
};

} // namespace phx
