// Micro-benchmark: how fast does a warp walk a long, fully unrolled FP64 instruction stream on sm_100a, as a
// function of resident warps per SM, and does lock-stepping the warps (barriers) change it?
#include <cstdio>
#include <cuda_runtime.h>
#ifndef BODY
#define BODY 8192
#endif
template <int CH, int SYNC_EVERY, bool LOOPED>
__global__ void __launch_bounds__(1024) k(double* out, double a, double b, int iters, long long* cyc) {
	extern __shared__ double sm[];
	double x[CH];
#pragma unroll
	for (int c = 0; c < CH; c++) x[c] = threadIdx.x * 1e-3 + c;
	long long t0 = clock64();
	if (LOOPED) {
#pragma unroll 1
		for (int it = 0; it < iters; it++) {
#pragma unroll
			for (int i = 0; i < 128 / CH; i++)
#pragma unroll
				for (int c = 0; c < CH; c++) x[c] = fma(x[c], a, b);
		}
	} else {
#pragma unroll
		for (int i = 0; i < BODY / CH; i++) {
#pragma unroll
			for (int c = 0; c < CH; c++) x[c] = fma(x[c], a, b);
			if (SYNC_EVERY > 0 && (i * CH) % SYNC_EVERY == 0) __syncthreads();
		}
	}
	long long t1 = clock64();
	double s = 0;
#pragma unroll
	for (int c = 0; c < CH; c++) s += x[c];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <typename K>
void run(const char* name, K kern, int block, int blocks_per_sm, int ninstr, int iters) {
	int smem = (200 * 1024) / blocks_per_sm;
	if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	int grid = 148 * blocks_per_sm;
	double* out; long long* cyc;
	cudaMalloc(&out, sizeof(double) * grid * block);
	cudaMalloc(&cyc, sizeof(long long) * grid);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int w = 0; w < 2; w++) kern<<<grid, block, smem>>>(out, 1.0000001, 1e-9, iters, cyc);
	cudaEventRecord(e0);
	kern<<<grid, block, smem>>>(out, 1.0000001, 1e-9, iters, cyc);
	cudaEventRecord(e1); cudaDeviceSynchronize();
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	long long h[148 * 32]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
	double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
	int warps_sm = block / 32 * blocks_per_sm;
	double ipc_smsp = (double)ninstr * warps_sm / 4.0 / avg;
	printf("%-28s block %4d x %d/SM = %2d warps/SM: %8.0f cyc/warp  %6.3f cyc/instr/warp  IPC/SMSP %.3f  fp64 pipe %.1f%%  (%.1f us) err=%s\n", name, block,
		   blocks_per_sm, warps_sm, avg, avg / ninstr, ipc_smsp, 200.0 * ipc_smsp, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
	cudaFree(out); cudaFree(cyc);
}
int main() {
	for (int bps : {1, 2, 4}) run("straight CH=4", k<4, 0, false>, 128, bps, BODY, 0);
	run("straight CH=4", k<4, 0, false>, 256, 4, BODY, 0);
	for (int bps : {1, 2, 4}) run("straight CH=8", k<8, 0, false>, 128, bps, BODY, 0);
	run("straight CH=8", k<8, 0, false>, 256, 4, BODY, 0);
	run("straight CH=1", k<1, 0, false>, 128, 1, BODY, 0);
	run("straight CH=2", k<2, 0, false>, 128, 2, BODY, 0);
	for (int bps : {1, 2, 4}) run("looped CH=4", k<4, 0, true>, 128, bps, 128 * 64, 64);
	for (int bps : {1, 2, 4}) run("looped CH=8", k<8, 0, true>, 128, bps, 128 * 64, 64);
	run("looped CH=1", k<1, 0, true>, 128, 1, 128 * 64, 64);
	run("sync256 CH=4", k<4, 256, false>, 256, 1, BODY, 0);
	run("sync256 CH=4", k<4, 256, false>, 512, 1, BODY, 0);
	run("sync256 CH=4", k<4, 256, false>, 1024, 1, BODY, 0);
	run("sync64 CH=4", k<4, 64, false>, 256, 1, BODY, 0);
	run("sync64 CH=4", k<4, 64, false>, 512, 1, BODY, 0);
	run("sync1024 CH=4", k<4, 1024, false>, 256, 1, BODY, 0);
	run("sync1024 CH=4", k<4, 1024, false>, 256, 2, BODY, 0);
	return 0;
}
