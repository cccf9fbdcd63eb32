// Probe: how many independent branches of ONE CUDA graph run concurrently on a B200, against the same chains on N streams?
// Each branch is a chain of `depth` kernels of one CTA that spin for `us` microseconds (latency-bound stand-ins for the frame chain).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o graph_branches graph_branches.cu && ./graph_branches
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__global__ void spin(long long cycles, int *sink) {
	const long long t0 = clock64();
	while (clock64() - t0 < cycles) { }
	if (sink && threadIdx.x == 9999) *sink = 1;
}
int main() {
	const int depth = 10;
	const long long cycles = 100000;      // ~51 us at 1.965 GHz
	cudaStream_t main_s; CK(cudaStreamCreateWithFlags(&main_s, cudaStreamNonBlocking));
	for (int n : {1, 16, 32, 48, 64, 96, 128}) {
		std::vector<cudaStream_t> st(n);
		std::vector<cudaEvent_t> done(n);
		for (int i = 0; i < n; i++) { CK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming)); }
		cudaEvent_t fork, e0, e1; CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming)); CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
		auto enqueue = [&]() {
			CK(cudaEventRecord(fork, main_s));
			for (int i = 0; i < n; i++) {
				CK(cudaStreamWaitEvent(st[i], fork, 0));
				for (int d = 0; d < depth; d++) spin<<<1, 32, 0, st[i]>>>(cycles, nullptr);
				CK(cudaEventRecord(done[i], st[i]));
				CK(cudaStreamWaitEvent(main_s, done[i], 0));
			}
		};
		// (a) plain streams
		enqueue(); CK(cudaStreamSynchronize(main_s));
		CK(cudaEventRecord(e0, main_s)); enqueue(); CK(cudaEventRecord(e1, main_s)); CK(cudaStreamSynchronize(main_s));
		float ms_streams = 0; CK(cudaEventElapsedTime(&ms_streams, e0, e1));
		// (b) the same work captured into one graph
		cudaGraph_t g; cudaGraphExec_t ge;
		CK(cudaStreamBeginCapture(main_s, cudaStreamCaptureModeGlobal));
		enqueue();
		CK(cudaStreamEndCapture(main_s, &g));
		CK(cudaGraphInstantiate(&ge, g, 0));
		CK(cudaGraphLaunch(ge, main_s)); CK(cudaStreamSynchronize(main_s));
		CK(cudaEventRecord(e0, main_s)); CK(cudaGraphLaunch(ge, main_s)); CK(cudaEventRecord(e1, main_s)); CK(cudaStreamSynchronize(main_s));
		float ms_graph = 0; CK(cudaEventElapsedTime(&ms_graph, e0, e1));
		printf("{\"branches\": %d, \"chain_kernels\": %d, \"ideal_ms\": %.3f, \"streams_ms\": %.3f, \"one_graph_ms\": %.3f}\n", n, depth, depth * cycles / 1.965e6, ms_streams, ms_graph);
		CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
		for (int i = 0; i < n; i++) { cudaStreamDestroy(st[i]); cudaEventDestroy(done[i]); }
	}
	return 0;
}
