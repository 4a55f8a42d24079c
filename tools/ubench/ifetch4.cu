// Micro-benchmark 4: where is the fast instruction-fetch window?  Straight-line FP64 segments of 16 KB, optionally
// preceded by / interleaved with cold (never executed) code, and a rolled loop placed late in the kernel.
#include <cstdio>
#include <cuda_runtime.h>
#define SEG 1024
#define NSEG 10
#define HOT(x) _Pragma("unroll") for (int i = 0; i < SEG / 4; i++) { x[0] = fma(x[0], a, b); x[1] = fma(x[1], a, b); x[2] = fma(x[2], a, b); x[3] = fma(x[3], a, b); }
#define COLD(x, n) if (never) { _Pragma("unroll") for (int i = 0; i < (n) * SEG / 4; i++) { x[0] = fma(x[0], b, a); x[1] = fma(x[1], b, a); x[2] = fma(x[2], b, a); x[3] = fma(x[3], b, a); } }
#define STAMP(s) { long long t1; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1), "+d"(x[0]), "+d"(x[1]), "+d"(x[2]), "+d"(x[3]) :: "memory"); if (threadIdx.x == 0) cyc[blockIdx.x * 16 + (s)] = t1 - t0; t0 = t1; }
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, double a, double b, int never, long long* cyc) {
	double x[4] = {threadIdx.x * 1e-3, 1.0, 2.0, 3.0};
	long long t0; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0), "+d"(x[0]), "+d"(x[1]), "+d"(x[2]), "+d"(x[3]) :: "memory");
	if (MODE == 1) COLD(x, 3)  // 48 KB of cold code first
	HOT(x) STAMP(0) HOT(x) STAMP(1)
	if (MODE == 2) COLD(x, 3)  // cold code after 32 KB of hot code
	HOT(x) STAMP(2) HOT(x) STAMP(3) HOT(x) STAMP(4) HOT(x) STAMP(5) HOT(x) STAMP(6) HOT(x) STAMP(7)
	if (MODE == 3) {  // rolled loop far from the start: 8 x 1024 instructions through a 2 KB body
#pragma unroll 1
		for (int it = 0; it < 64; it++) {
#pragma unroll
			for (int i = 0; i < 32; i++) { x[0] = fma(x[0], a, b); x[1] = fma(x[1], a, b); x[2] = fma(x[2], a, b); x[3] = fma(x[3], a, b); }
		}
		STAMP(8)
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
}
__global__ void spin(long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
template <typename K>
void run(const char* name, K kern, int block, int bps, int nst) {
	int grid = 148 * bps, smem = 200 * 1024 / bps;
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	double* out; long long* cyc;
	cudaMalloc(&out, sizeof(double) * grid * block); cudaMalloc(&cyc, sizeof(long long) * grid * 16);
	for (int w = 0; w < 3; w++) kern<<<grid, block, smem>>>(out, 1.0000001, 1e-9, 0, cyc);
	cudaDeviceSynchronize();
	static long long h[148 * 8 * 16];
	cudaMemcpy(h, cyc, sizeof(long long) * grid * 16, cudaMemcpyDeviceToHost);
	printf("%-22s block %d x %d/SM: cyc/instr per 16KB segment:", name, block, bps);
	for (int s = 0; s < nst; s++) { double a = 0; for (int g = 0; g < grid; g++) a += h[g * 16 + s]; printf(" %.2f", a / grid / (s == 8 ? 8192 : SEG)); }
	printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
	cudaFree(out); cudaFree(cyc);
}
int main() {
	spin<<<148, 32>>>(400000000LL); cudaDeviceSynchronize();
	for (int bps : {1, 2}) {
		run("plain", k<0>, 128, bps, 8);
		run("48KB cold first", k<1>, 128, bps, 8);
		run("48KB cold after 32KB", k<2>, 128, bps, 8);
		run("rolled loop at end", k<3>, 128, bps, 9);
	}
	return 0;
}
