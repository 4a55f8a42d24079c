// Instantiations of the fused cycle kernel, one translation unit per robot dof (OSC_INST_N)
// so that nvcc compiles them in parallel.
#include "osc_cycle.cuh"
#include "osc_singular.cuh"
#include "osc_blend.cuh"
#include "osc_launch.h"

#ifndef OSC_INST_N
#error "compile with -DOSC_INST_N=<dof>"
#endif

// OSC_ONLY_R=<r> (tuning builds only): instantiate just the [MotionForceTask rank r, JointTask] kernel
#ifndef OSC_ONLY_R
#define OSC_ONLY_R 0
#endif

namespace osc {

template <int N, int R, bool JT, bool FULL>
static cudaError_t launch_variant(const OscProgram& P, cudaStream_t stream) {
	const unsigned grid = (unsigned)((P.n_robots + kCycleBlock - 1) / kCycleBlock);
	constexpr int smem = 9 * N * kCycleBlock * (int)sizeof(double);  // body orientations staged between the two passes
	static bool configured[64] = {false};  // per device: function attributes belong to the device's context
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64 || !configured[dev]) {
		cudaError_t e = cudaFuncSetAttribute(osc_cycle_kernel<N, R, JT, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		if (e != cudaSuccess) return e;
		if (dev >= 0 && dev < 64) configured[dev] = true;
	}
	osc_cycle_kernel<N, R, JT, FULL><<<grid, kCycleBlock, smem, stream>>>(P);
	return cudaGetLastError();
}

template <int N, int R, bool JT>
static cudaError_t launch_one(const OscProgram& P, cudaStream_t stream) {
	cudaError_t e0;
	if constexpr (R == 6) {
		e0 = (P.mft[0].full) ? launch_variant<N, R, JT, true>(P, stream) : launch_variant<N, R, JT, false>(P, stream);
	} else {
		e0 = launch_variant<N, R, JT, false>(P, stream);
	}
	if (e0 != cudaSuccess) return e0;
	if constexpr (R > 0) {
		// The general path for the robots the fast kernel handed over (usually none or few).  One block per SM,
		// grid-stride over the compacted list; launched with programmatic stream serialization so that its launch
		// latency overlaps the tail of the fast kernel (it waits on griddepcontrol.wait before reading the list).
		const long long want = (P.n_robots + 63) / 64;
		cudaLaunchConfig_t cfg{};
		cfg.gridDim = dim3((unsigned)(want < 148 * OSC_GENERIC_MIN_BLOCKS ? want : 148 * OSC_GENERIC_MIN_BLOCKS));
		cfg.blockDim = dim3(64);
		cfg.dynamicSmemBytes = 0;
		cfg.stream = stream;
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = attr;
		cfg.numAttrs = 1;
		if constexpr (R == 6) {
			// flagship hierarchy: unrolled blending path (osc_blend.cuh), general path only as its fallback
			if (P.mft[0].full) return cudaLaunchKernelEx(&cfg, osc_blend_kernel<N, JT>, P);
		}
		return cudaLaunchKernelEx(&cfg, osc_singular_kernel<N>, P);
	}
	return cudaGetLastError();
}

// motion-force task of rank R (1..min(6,N)), with or without a closing full joint task
template <int N, int R>
static cudaError_t launch_rank(bool has_jt, const OscProgram& P, cudaStream_t stream) {
	if constexpr (R >= 1 && R <= N && R <= 6 && OSC_ONLY_R != 0) {
		if constexpr (R == OSC_ONLY_R) return has_jt ? launch_one<N, R, true>(P, stream) : cudaErrorNotSupported;
		return cudaErrorNotSupported;
	} else if constexpr (R >= 1 && R <= N && R <= 6) {
		return has_jt ? launch_one<N, R, true>(P, stream) : launch_one<N, R, false>(P, stream);
	} else {
		return cudaErrorNotSupported;
	}
}

template <int N>
static cudaError_t launch_generic(const OscProgram& P, cudaStream_t stream) {
	const unsigned grid = (unsigned)((P.n_robots + 63) / 64);
	osc_generic_kernel<N><<<grid, 64, 0, stream>>>(P);
	return cudaGetLastError();
}

#define CONCAT_(a, b) a##b
#define CONCAT(a, b) CONCAT_(a, b)

cudaError_t CONCAT(launch_cycle_n, OSC_INST_N)(int R, bool has_jt, const OscProgram& P, cudaStream_t stream) {
	constexpr int N = OSC_INST_N;
	if (R < 0) return launch_generic<N>(P, stream);	 // general hierarchy: every robot on the general path
	switch (R) {
		case 0:
			if constexpr (OSC_ONLY_R != 0) return cudaErrorNotSupported;
			else return has_jt ? launch_one<N, 0, true>(P, stream) : cudaErrorNotSupported;
		case 1: return launch_rank<N, 1>(has_jt, P, stream);
		case 2: return launch_rank<N, 2>(has_jt, P, stream);
		case 3: return launch_rank<N, 3>(has_jt, P, stream);
		case 4: return launch_rank<N, 4>(has_jt, P, stream);
		case 5: return launch_rank<N, 5>(has_jt, P, stream);
		case 6: return launch_rank<N, 6>(has_jt, P, stream);
	}
	return cudaErrorNotSupported;
}

}  // namespace osc
