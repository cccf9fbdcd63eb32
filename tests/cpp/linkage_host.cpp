// tests/cpp/linkage_host.cpp — TEST INFRASTRUCTURE: the device source of the cached-row-maximum agglomeration
// (moped_b200/csrc/linkage_cached.cuh) compiled by g++ with the threads of the block emulated by a loop, so that its merge
// sequence can be checked against the oracle on a machine without a GPU (tests/test_linkage_cached_host.py). Built with
// -ffp-contract=off (linkage.cu is built with -fmad=false). Not part of the product.
#include "../../moped_b200/csrc/linkage_cached.cuh"

#include <vector>

extern "C" void lh_agglomerate(int thread_order, int n, const float *K, float cutoff, int min_pts, int *out_count, int *out_offsets,
                               int *out_members) {
	constexpr int W = 256;                                    // kLinkThreads
	const size_t nn = (size_t)n * n;
	std::vector<float> D(nn), f3(3 * (size_t)n), s_val(W);
	std::vector<int> L(n), lists(nn), sz(n), i3(3 * (size_t)n), s_idx(W), ctl(16);
	lkx::State s;
	s.n = n; s.D = D.data(); s.L = L.data(); s.lists = lists.data(); s.sz = sz.data();
	s.tmp = f3.data(); s.oldcol = f3.data() + n; s.best = f3.data() + 2 * (size_t)n;
	s.posOf = i3.data(); s.arg = i3.data() + n; s.tmpi = i3.data() + 2 * (size_t)n;
	s.s_val = s_val.data(); s.s_idx = s_idx.data(); s.ctl = ctl.data();
	phx::g_host_order = thread_order;
	lkx::Block<W> blk;
	blk.tid = 0;
	lkx::agglomerate_average(blk, s, K, cutoff, min_pts, out_count, out_offsets, out_members);
}
