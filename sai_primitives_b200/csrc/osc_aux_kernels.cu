#include "osc_aux_kernels.cuh"
#include "osc_launch.h"

namespace osc {

static inline dim3 grid_for(int64_t n, int block) { return dim3((unsigned)((n + block - 1) / block)); }

#define DISPATCH_N(n, CALL)                  \
	switch (n) {                             \
		case 1: { constexpr int N_ = 1; CALL; } break; \
		case 2: { constexpr int N_ = 2; CALL; } break; \
		case 3: { constexpr int N_ = 3; CALL; } break; \
		case 4: { constexpr int N_ = 4; CALL; } break; \
		case 5: { constexpr int N_ = 5; CALL; } break; \
		case 6: { constexpr int N_ = 6; CALL; } break; \
		case 7: { constexpr int N_ = 7; CALL; } break; \
		case 8: { constexpr int N_ = 8; CALL; } break; \
		default: return cudaErrorInvalidValue;         \
	}

cudaError_t launch_reinit_mft(const OscProgram& P, int mft_index, int full_init, cudaStream_t stream) {
	DISPATCH_N(P.model.n, (reinit_mft_kernel<N_><<<grid_for(P.n_robots, 128), 128, 0, stream>>>(P, mft_index, full_init)));
	return cudaGetLastError();
}
cudaError_t launch_sim_integrate(const OscProgram& P, double* q, double* dq, const double* tau, double dt, int substeps, cudaStream_t stream) {
	DISPATCH_N(P.model.n, (sim_integrate_kernel<N_><<<grid_for(P.n_robots, 128), 128, 0, stream>>>(P, q, dq, tau, dt, substeps)));
	return cudaGetLastError();
}
cudaError_t launch_jla(const OscProgram& P, cudaStream_t stream) {
	DevJla jp;	// JointLimitAvoidanceTask.h:26-35
	jp.kv = 20.0;
	jp.position_z1_to_limit = 9.0 * 3.14159265358979323846 / 180.0;
	jp.position_z2_to_limit = 6.0 * 3.14159265358979323846 / 180.0;
	jp.velocity_z1_to_limit = 0.5;
	jp.velocity_z2_to_limit = 0.3;
	jp.max_torque_ratio_pos_limit = 1.0;
	jp.max_torque_ratio_vel_limit = 0.05;
	DISPATCH_N(P.model.n, (jla_kernel<N_><<<grid_for(P.n_robots, 128), 128, 0, stream>>>(P, jp)));
	return cudaGetLastError();
}
cudaError_t launch_reinit_jt(const OscProgram& P, int jt_index, cudaStream_t stream) {
	DISPATCH_N(P.model.n, (reinit_jt_kernel<N_><<<grid_for(P.n_robots, 128), 128, 0, stream>>>(P, jt_index)));
	return cudaGetLastError();
}
cudaError_t launch_sensed_wrench(const OscProgram& P, int mft_index, const double* f, const double* m, cudaStream_t stream) {
	DISPATCH_N(P.model.n, (sensed_wrench_kernel<N_><<<grid_for(P.n_robots, 128), 128, 0, stream>>>(P, mft_index, f, m)));
	return cudaGetLastError();
}
cudaError_t launch_eval_model(const OscProgram& P, const osc_link_frame& frame, double* M, double* J, double* x, double* R,
							  double* g, cudaStream_t stream) {
	EvalOut out{M, J, x, R, g};
	DISPATCH_N(P.model.n, (eval_model_kernel<N_><<<grid_for(P.n_robots, 128), 128, 0, stream>>>(P, frame, out)));
	return cudaGetLastError();
}
cudaError_t launch_fill(double* st, int64_t NR, int comp, int ncomp, const double* vals, cudaStream_t stream) {
	if (ncomp > 16) return cudaErrorInvalidValue;
	FillVals v;
	for (int c = 0; c < 16; c++) v.v[c] = (c < ncomp) ? vals[c] : 0.0;
	fill_components_kernel<<<grid_for(NR, 256), 256, 0, stream>>>(st, NR, comp, ncomp, v);
	return cudaGetLastError();
}
cudaError_t launch_copy(double* st, int64_t NR, int dst, int src, int ncomp, cudaStream_t stream) {
	copy_components_kernel<<<grid_for(NR, 256), 256, 0, stream>>>(st, NR, dst, src, ncomp);
	return cudaGetLastError();
}
cudaError_t launch_fill_int(int32_t* ist, int64_t NR, int comp, int ncomp, int32_t value, cudaStream_t stream) {
	fill_int_components_kernel<<<grid_for(NR, 256), 256, 0, stream>>>(ist, NR, comp, ncomp, value);
	return cudaGetLastError();
}

}  // namespace osc
