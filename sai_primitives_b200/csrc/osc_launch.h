// Host-callable launch wrappers around the templated kernels (implemented in the .cu files).
#pragma once
#include <cuda_runtime.h>

#include "osc_dev_types.h"

namespace osc {

#ifndef OSC_CYCLE_BLOCK
#define OSC_CYCLE_BLOCK 128
#endif
constexpr int kCycleBlock = OSC_CYCLE_BLOCK;

// robot dofs the fused cycle kernel is instantiated for (one translation unit each, see Makefile)
#ifdef OSC_TUNE_DOF7_ONLY  // tuning builds (tools/tune.sh): the Panda kernels only
#define OSC_CYCLE_DOFS(X) X(7)
#else
#define OSC_CYCLE_DOFS(X) X(4) X(6) X(7) X(8)
#endif

// Fused control cycle for hierarchy signature (n, R, has_jt): R = rank of a leading MotionForceTask (0: none),
// has_jt = a full JointTask closes the hierarchy; R < 0 selects the general-hierarchy kernel.
// cudaErrorNotSupported when the dof is not compiled in (OSC_CYCLE_DOFS).
cudaError_t launch_cycle(int n, int R, bool has_jt, const OscProgram& P, cudaStream_t stream);
bool cycle_signature_available(int n, int R, bool has_jt);
// the general path of this launch runs as the three kernels of the split blending path (osc_blend.cuh) instead of osc_blend_kernel:
// flagship hierarchy, pipelined handle, and the host hint says that many robots are on the general path (osc_capi.cu, run_cycle)
inline bool blend_split_selected(const OscProgram& P) {
	return P.blend_split_on && P.blend_scratch != nullptr && P.general_done != nullptr && !P.general_grid_small && P.n_tasks >= 1 && P.tasks[0].type == OSC_TASK_MOTION_FORCE &&
		   P.mft[0].full && P.mft[0].rank == 6;
}

cudaError_t launch_popc_probe(const OscProgram& P, int mft_index, int K, const double* fd, const double* fs, const double* vcl, const double* vr,
							  double kv, double kff, double* out, cudaStream_t stream);
cudaError_t measure_fp64_peak(double seconds, double* tflops, cudaStream_t stream);
cudaError_t launch_sim_integrate(const OscProgram& P, double* q, double* dq, const double* tau, double dt, int substeps, cudaStream_t stream);
cudaError_t launch_jla(const OscProgram& P, cudaStream_t stream);
// internal OTG of the tasks (osc_otg_kernels.cuh): per-cycle update, (re)initialisation (mode 0 whole task, 1 linear part,
// 2 angular part, 3 limits changed; fresh = generator constructed), switch-off
cudaError_t launch_otg_update(const OscProgram& P, cudaStream_t stream);
cudaError_t launch_otg_init(const OscProgram& P, int task, int mode, int fresh, cudaStream_t stream);
cudaError_t launch_otg_disable(const OscProgram& P, int task, cudaStream_t stream);
// read-only observers (osc_observers.cuh): kind = ObserverKind, task = position in the hierarchy, out = ncomp x N (SoA)
cudaError_t launch_observer(const OscProgram& P, int task, int kind, double* out, cudaStream_t stream);
cudaError_t launch_reinit_mft(const OscProgram& P, int mft_index, int full_init, cudaStream_t stream);
cudaError_t launch_reinit_jt(const OscProgram& P, int jt_index, cudaStream_t stream);
cudaError_t launch_sensed_wrench(const OscProgram& P, int mft_index, const double* f, const double* m, cudaStream_t stream);
cudaError_t launch_eval_model(const OscProgram& P, const osc_link_frame& frame, double* M, double* J, double* x, double* R,
							  double* g, cudaStream_t stream);
cudaError_t launch_fill(double* st, int64_t NR, int comp, int ncomp, const double* vals, cudaStream_t stream);
cudaError_t launch_copy(double* st, int64_t NR, int dst, int src, int ncomp, cudaStream_t stream);
cudaError_t launch_fill_int(int32_t* ist, int64_t NR, int comp, int ncomp, int32_t value, cudaStream_t stream);

}  // namespace osc
