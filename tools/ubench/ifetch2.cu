// Micro-benchmark 2: straight-line FP64 stream with warps de-synchronised by a start skew; clock calibration.
#include <cstdio>
#include <cuda_runtime.h>
#ifndef BODY
#define BODY 8192
#endif
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void spin(long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
template <int CH, int SYNC_EVERY>
__global__ void __launch_bounds__(1024) k(double* out, double a, double b, int skew, long long* cyc, unsigned long long* ns) {
	double x[CH];
#pragma unroll
	for (int c = 0; c < CH; c++) x[c] = threadIdx.x * 1e-3 + c;
	if (skew > 0) {
		unsigned w = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2654435761u;
		long long d = (long long)((w >> 20) & 1023) * skew / 1024;
		long long s0 = clock64();
		while (clock64() - s0 < d) {}
	}
	long long t0 = clock64();
	unsigned long long g0 = gtime();
#pragma unroll
	for (int i = 0; i < BODY / CH; i++) {
#pragma unroll
		for (int c = 0; c < CH; c++) x[c] = fma(x[c], a, b);
		if (SYNC_EVERY > 0 && (i * CH) % SYNC_EVERY == 0) __syncthreads();
	}
	long long t1 = clock64();
	unsigned long long g1 = gtime();
	double s = 0;
#pragma unroll
	for (int c = 0; c < CH; c++) s += x[c];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if ((threadIdx.x & 31) == 0) { int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); cyc[w] = t1 - t0; ns[w] = g1 - g0; }
}
template <typename K>
void run(const char* name, K kern, int block, int blocks_per_sm, int skew) {
	int smem = (200 * 1024) / blocks_per_sm;
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	int grid = 148 * blocks_per_sm, nw = grid * block / 32;
	double* out; long long* cyc; unsigned long long* ns;
	cudaMalloc(&out, sizeof(double) * grid * block);
	cudaMalloc(&cyc, sizeof(long long) * nw); cudaMalloc(&ns, sizeof(long long) * nw);
	for (int w = 0; w < 3; w++) kern<<<grid, block, smem>>>(out, 1.0000001, 1e-9, skew, cyc, ns);
	cudaDeviceSynchronize();
	static long long h[148 * 64]; static unsigned long long hn[148 * 64];
	cudaMemcpy(h, cyc, sizeof(long long) * nw, cudaMemcpyDeviceToHost); cudaMemcpy(hn, ns, sizeof(long long) * nw, cudaMemcpyDeviceToHost);
	double avg = 0, an = 0; for (int i = 0; i < nw; i++) { avg += h[i]; an += hn[i]; } avg /= nw; an /= nw;
	int warps_sm = block / 32 * blocks_per_sm;
	printf("%-16s block %4d x %d/SM = %2d warps/SM skew %6d: %8.0f cyc/warp %6.3f cyc/instr/warp  IPC/SMSP %.3f   clock %.3f GHz  err=%s\n", name, block, blocks_per_sm,
		   warps_sm, skew, avg, avg / BODY, (double)BODY * warps_sm / 4.0 / avg, avg / an, cudaGetErrorString(cudaGetLastError()));
	cudaFree(out); cudaFree(cyc); cudaFree(ns);
}
int main() {
	spin<<<148, 32>>>(400000000LL); cudaDeviceSynchronize();  // ~0.2 s to let the clocks ramp
	for (int skew : {0, 2000, 20000, 100000}) {
		run("straight CH=4", k<4, 0>, 128, 1, skew);
		run("straight CH=4", k<4, 0>, 128, 2, skew);
		run("straight CH=4", k<4, 0>, 128, 4, skew);
		run("straight CH=4", k<4, 0>, 256, 1, skew);
	}
	for (int skew : {20000}) {
		run("sync256 CH=4", k<4, 256>, 256, 1, skew);
		run("sync256 CH=4", k<4, 256>, 128, 2, skew);
		run("sync1024 CH=4", k<4, 1024>, 256, 1, skew);
		run("sync1024 CH=4", k<4, 1024>, 512, 1, skew);
	}
	return 0;
}
