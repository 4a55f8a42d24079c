// Micro-benchmark 3: per-segment fetch cost along a long straight-line FP64 stream, with and without a taken jump
// over cold code in the middle (does the sequential instruction prefetch survive a long forward branch?).
#include <cstdio>
#include <cuda_runtime.h>
#define SEG 1024  // instructions per timed segment (16 KB)
#define NSEG 10
template <int JUMP_AFTER>  // segment index after which a forward jump over dead code is taken (-1: none)
__global__ void __launch_bounds__(256) k(double* out, double a, double b, int never, long long* cyc) {
	double x[4];
#pragma unroll
	for (int c = 0; c < 4; c++) x[c] = threadIdx.x * 1e-3 + c;
	long long t[NSEG + 1];
	t[0] = clock64();
#pragma unroll
	for (int s = 0; s < NSEG; s++) {
#pragma unroll
		for (int i = 0; i < SEG / 4; i++)
#pragma unroll
			for (int c = 0; c < 4; c++) x[c] = fma(x[c], a, b);
		t[s + 1] = clock64();
		if (s == JUMP_AFTER) {
			if (never) {  // 3 segments (48 KB) of code that is never executed
#pragma unroll
				for (int i = 0; i < 3 * SEG / 4; i++)
#pragma unroll
					for (int c = 0; c < 4; c++) x[c] = fma(x[c], b, a);
			}
		}
	}
	double sum = 0;
#pragma unroll
	for (int c = 0; c < 4; c++) sum += x[c];
	out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
	if (threadIdx.x == 0)
		for (int s = 0; s < NSEG; s++) cyc[blockIdx.x * NSEG + s] = t[s + 1] - t[s];
}
__global__ void spin(long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
template <typename K>
void run(const char* name, K kern, int block, int bps) {
	int grid = 148 * bps, smem = 200 * 1024 / bps;
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	double* out; long long* cyc;
	cudaMalloc(&out, sizeof(double) * grid * block); cudaMalloc(&cyc, sizeof(long long) * grid * NSEG);
	for (int w = 0; w < 3; w++) kern<<<grid, block, smem>>>(out, 1.0000001, 1e-9, 0, cyc);
	cudaDeviceSynchronize();
	static long long h[148 * 8 * NSEG];
	cudaMemcpy(h, cyc, sizeof(long long) * grid * NSEG, cudaMemcpyDeviceToHost);
	printf("%-10s block %d x %d/SM: cyc/instr per 16KB segment:", name, block, bps);
	for (int s = 0; s < NSEG; s++) { double a = 0; for (int g = 0; g < grid; g++) a += h[g * NSEG + s]; printf(" %.2f", a / grid / SEG); }
	printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
	cudaFree(out); cudaFree(cyc);
}
int main() {
	spin<<<148, 32>>>(400000000LL); cudaDeviceSynchronize();
	run("nojump", k<-1>, 128, 1); run("nojump", k<-1>, 128, 2); run("nojump", k<-1>, 256, 1);
	run("jump@3", k<2>, 128, 1); run("jump@3", k<2>, 128, 2); run("jump@3", k<2>, 256, 1);
	return 0;
}
