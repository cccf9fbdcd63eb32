// sm_partition.cu — spatial partition of the GPU between MATCH and the stages after it.
//
// The coarse matching kernel is persistent (one CTA per SM, all of the SM's shared memory); the stage chains CLUSTER..FILTER2
// of the frames are thousands of small, latency-bound CTAs. Left to the hardware scheduler they do not overlap: a stage CTA on
// an SM keeps a whole matching CTA out (measured: MATCH + stages of the previous batch enqueued together take as long as one
// after the other). CUDA green contexts give each side its own SMs: `stage_sms` SMs (a multiple of 8 on sm_100) for the
// streams of the frame lanes, the rest for the stream the coarse kernel is launched on.
//
// The driver entry points are fetched with cudaGetDriverEntryPoint, so libmoped_cuda.so keeps loading on machines without
// libcuda (the CPU-side symbol tests).
#include "common.cuh"

#include <cuda.h>

namespace mc {

namespace {
struct DriverApi {
	CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
	CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
	CUresult (*DevSmResourceSplitByCount)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
	CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
	CUresult (*GreenCtxCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
	CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
	CUresult (*GreenCtxStreamCreate)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
	bool ok = false;
};

bool load_driver(DriverApi &d, std::string &err) {
	struct { const char *name; void **fn; } want[] = {
		{ "cuDeviceGet", (void **)&d.DeviceGet }, { "cuDeviceGetDevResource", (void **)&d.DeviceGetDevResource },
		{ "cuDevSmResourceSplitByCount", (void **)&d.DevSmResourceSplitByCount }, { "cuDevResourceGenerateDesc", (void **)&d.DevResourceGenerateDesc },
		{ "cuGreenCtxCreate", (void **)&d.GreenCtxCreate }, { "cuGreenCtxDestroy", (void **)&d.GreenCtxDestroy },
		{ "cuGreenCtxStreamCreate", (void **)&d.GreenCtxStreamCreate } };
	for (auto &w : want) {
		cudaDriverEntryPointQueryResult st;
		if (cudaGetDriverEntryPoint(w.name, w.fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !*w.fn) {
			cudaGetLastError();
			err = std::string("sm partition: driver entry point ") + w.name + " not available";
			return false;
		}
	}
	d.ok = true;
	return true;
}
} // namespace

void sm_partition_destroy(mc_ctx *ctx) {
	if (!ctx->green_match && !ctx->green_stage) return;
	DriverApi d;
	std::string err;
	if (ctx->match_green_stream) { cudaStreamSynchronize(ctx->match_green_stream); cudaStreamDestroy(ctx->match_green_stream); ctx->match_green_stream = nullptr; }
	if (ctx->ev_green[0]) { cudaEventDestroy(ctx->ev_green[0]); cudaEventDestroy(ctx->ev_green[1]); ctx->ev_green[0] = ctx->ev_green[1] = nullptr; }
	if (load_driver(d, err)) {
		if (ctx->green_match) d.GreenCtxDestroy((CUgreenCtx)ctx->green_match);
		if (ctx->green_stage) d.GreenCtxDestroy((CUgreenCtx)ctx->green_stage);
	}
	ctx->green_match = ctx->green_stage = nullptr;
	ctx->match_sms = ctx->stage_sms = 0;
}

// stage_sms = 0 removes the partition. The lanes are destroyed by the caller first: their streams belong to the old partition.
mc_status sm_partition_create(mc_ctx *ctx, int stage_sms) {
	sm_partition_destroy(ctx);
	if (stage_sms <= 0) return MC_OK;
	DriverApi d;
	if (!load_driver(d, ctx->err)) return MC_ERR_STATE;
	auto fail = [&](const char *what, CUresult r) { ctx->err = std::string("sm partition: ") + what + " failed (CUresult " + std::to_string((int)r) + ")"; return MC_ERR_CUDA; };
	CUdevice dev;
	CUresult r = d.DeviceGet(&dev, ctx->device);
	if (r != CUDA_SUCCESS) return fail("cuDeviceGet", r);
	CUdevResource all, grp, rem;
	r = d.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM);
	if (r != CUDA_SUCCESS) return fail("cuDeviceGetDevResource", r);
	unsigned int nb = 1;
	r = d.DevSmResourceSplitByCount(&grp, &nb, &all, &rem, 0, (unsigned)stage_sms);
	if (r != CUDA_SUCCESS || nb != 1) return fail("cuDevSmResourceSplitByCount", r);
	if (rem.sm.smCount < 16) { ctx->err = "sm partition: fewer than 16 SMs would be left for MATCH"; return MC_ERR_ARG; }
	CUdevResourceDesc desc_stage, desc_match;
	r = d.DevResourceGenerateDesc(&desc_stage, &grp, 1);
	if (r != CUDA_SUCCESS) return fail("cuDevResourceGenerateDesc", r);
	r = d.DevResourceGenerateDesc(&desc_match, &rem, 1);
	if (r != CUDA_SUCCESS) return fail("cuDevResourceGenerateDesc", r);
	CUgreenCtx g_stage = nullptr, g_match = nullptr;
	r = d.GreenCtxCreate(&g_stage, desc_stage, dev, CU_GREEN_CTX_DEFAULT_STREAM);
	if (r != CUDA_SUCCESS) return fail("cuGreenCtxCreate", r);
	r = d.GreenCtxCreate(&g_match, desc_match, dev, CU_GREEN_CTX_DEFAULT_STREAM);
	if (r != CUDA_SUCCESS) { d.GreenCtxDestroy(g_stage); return fail("cuGreenCtxCreate", r); }
	CUstream ms = nullptr;
	r = d.GreenCtxStreamCreate(&ms, g_match, CU_STREAM_NON_BLOCKING, 0);
	if (r != CUDA_SUCCESS) { d.GreenCtxDestroy(g_stage); d.GreenCtxDestroy(g_match); return fail("cuGreenCtxStreamCreate", r); }
	ctx->green_stage = g_stage; ctx->green_match = g_match;
	ctx->match_green_stream = (cudaStream_t)ms;
	ctx->stage_sms = (int)grp.sm.smCount; ctx->match_sms = (int)rem.sm.smCount;
	MC_CUDA(cudaEventCreateWithFlags(&ctx->ev_green[0], cudaEventDisableTiming));
	MC_CUDA(cudaEventCreateWithFlags(&ctx->ev_green[1], cudaEventDisableTiming));
	return MC_OK;
}

// a stream confined to the stage partition (frame lanes)
mc_status sm_partition_stage_stream(mc_ctx *ctx, cudaStream_t *out, int priority) {
	DriverApi d;
	if (!load_driver(d, ctx->err)) return MC_ERR_STATE;
	CUstream s = nullptr;
	CUresult r = d.GreenCtxStreamCreate(&s, (CUgreenCtx)ctx->green_stage, CU_STREAM_NON_BLOCKING, priority);
	if (r != CUDA_SUCCESS) { ctx->err = "sm partition: cuGreenCtxStreamCreate failed (CUresult " + std::to_string((int)r) + ")"; return MC_ERR_CUDA; }
	*out = (cudaStream_t)s;
	return MC_OK;
}

} // namespace mc
