// Small per-robot service kernels around the fused cycle: task (re)initialisation from the
// current state, sensed-wrench frame change, the stand-alone kinematics/dynamics export
// used by the parity tests, and SoA component fills/copies.
#pragma once
#include "osc_kindyn.cuh"
#include "osc_tasks.cuh"

namespace osc {

#define ST(comp, c) st[(int64_t)((comp) + (c)) * NR + i]

// value fill of `ncomp` consecutive SoA components (vals in kernel parameter space)
struct FillVals {
	double v[16];
};
__global__ void fill_components_kernel(double* st, int64_t NR, int comp, int ncomp, FillVals vals) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= NR) return;
	for (int c = 0; c < ncomp; c++) ST(comp, c) = vals.v[c];
}
__global__ void copy_components_kernel(double* st, int64_t NR, int dst, int src, int ncomp) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= NR) return;
	for (int c = 0; c < ncomp; c++) ST(dst, c) = ST(src, c);
}
__global__ void fill_int_components_kernel(int32_t* ist, int64_t NR, int comp, int ncomp, int32_t value) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= NR) return;
	for (int c = 0; c < ncomp; c++) ist[(int64_t)(comp + c) * NR + i] = value;
}

// MotionForceTask::reInitializeTask (MotionForceTask.cpp:204-245) + the singularity handler /
// POPC initial state (SingularityHandler.cpp:53-63, POPCExplicitForceControl.cpp:10-21).
// full_init = 1 also resets what only the constructor resets (handler + POPC state).
template <int N>
__global__ void reinit_mft_kernel(const __grid_constant__ OscProgram P, int mft_index, int full_init) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t NR = P.n_robots;
	if (i >= NR) return;
	const DevMft& t = P.mft[mft_index];
	double* st = t.st;
	double q[N];
#pragma unroll
	for (int j = 0; j < N; j++) q[j] = P.q[(int64_t)j * NR + i];
	KinDyn<N> kd;
	forward_kinematics<N>(P.model, q, kd);
	double x[3], R[9];
	frame_pose<N>(kd, t.body, t.ctrl_R, t.ctrl_t, x, R);
	for (int c = 0; c < MC_POPC; c++) ST(0, c) = 0.0;  // goals, sensed values, integrators, observers
	store3(st, NR, i, MC_GOAL_POS, x);
	store3(st, NR, i, MC_CUR_POS, x);
#pragma unroll
	for (int k = 0; k < 9; k++) {
		ST(MC_GOAL_ORI, k) = R[k];
		ST(MC_CUR_ORI, k) = R[k];
	}
	if (full_init) {
		ST(MC_POPC, 0) = 0.0;
		ST(MC_POPC, 1) = 0.0;
		ST(MC_POPC, 2) = 1.0;  // Rc
		ST(MC_POPC, 3) = 0.0;
		for (int j = 0; j < OSC_MAX_DOF; j++) {
			const double mid = (j < N) ? 0.5 * (P.model.q_lower[j] + P.model.q_upper[j]) : 0.0;
			ST(MC_Q_PRIOR, j) = mid;
			ST(MC_DQ_PRIOR, j) = 0.0;
			ST(MC_TYPE2_DIR, j) = 1.0;
		}
		for (int c = 0; c < MI_COUNT; c++) t.ist[(int64_t)c * NR + i] = 0;
		t.ist[(int64_t)MI_POPC_COUNTER * NR + i] = 50;
	}
}

// JointTask::reInitializeTask (JointTask.cpp:91-107)
template <int N>
__global__ void reinit_jt_kernel(const __grid_constant__ OscProgram P, int jt_index) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t NR = P.n_robots;
	if (i >= NR) return;
	const DevJt& t = P.jt[jt_index];
	double* st = t.st;
	for (int c = 0; c < JC_COUNT; c++) ST(0, c) = 0.0;
	for (int a = 0; a < t.k; a++) {
		double pos = 0.0;
#pragma unroll
		for (int j = 0; j < N; j++) pos += t.S[a][j] * P.q[(int64_t)j * NR + i];
		ST(JC_GOAL_POS, a) = pos;
	}
}

// MotionForceTask::updateSensedForceAndMoment (MotionForceTask.cpp:805-828)
template <int N>
__global__ void sensed_wrench_kernel(const __grid_constant__ OscProgram P, int mft_index, const double* f_sensor,
									 const double* m_sensor) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t NR = P.n_robots;
	if (i >= NR) return;
	const DevMft& t = P.mft[mft_index];
	double* st = t.st;
	double q[N];
#pragma unroll
	for (int j = 0; j < N; j++) q[j] = P.q[(int64_t)j * NR + i];
	KinDyn<N> kd;
	forward_kinematics<N>(P.model, q, kd);
	double x[3], R[9];
	frame_pose<N>(kd, t.body, t.ctrl_R, t.ctrl_t, x, R);
	double fs[3], ms[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		fs[k] = f_sensor[(int64_t)k * NR + i];
		ms[k] = m_sensor[(int64_t)k * NR + i];
	}
	store3(st, NR, i, MC_SENSED_F_SENSOR, fs);
	store3(st, NR, i, MC_SENSED_M_SENSOR, ms);
	double f[3], m[3], tmp[3], cr[3];
	mat3_vec(t.cs_R, fs, f);
	mat3_vec(t.cs_R, ms, tmp);
	cross3(t.cs_t, f, cr);
#pragma unroll
	for (int k = 0; k < 3; k++) m[k] = cr[k] + tmp[k];
	double fw[3], mw[3];
	mat3_vec(R, f, fw);
	mat3_vec(R, m, mw);
	store3(st, NR, i, MC_SENSED_F, fw);
	store3(st, NR, i, MC_SENSED_M, mw);
}

// Kinematics/dynamics stage on its own: M, J (6 x n, linear first), x, R, g.
struct EvalOut {
	double *M, *J, *x, *R, *g;
};
template <int N>
__global__ void eval_model_kernel(const __grid_constant__ OscProgram P, osc_link_frame frame, EvalOut out) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t NR = P.n_robots;
	if (i >= NR) return;
	double q[N];
#pragma unroll
	for (int j = 0; j < N; j++) q[j] = P.q[(int64_t)j * NR + i];
	KinDyn<N> kd;
	forward_kinematics<N>(P.model, q, kd);
	mass_matrix<N, true>(P.model, kd);
	double x[3], R[9];
	frame_pose<N>(kd, frame.body, frame.R, frame.t, x, R);
	double JT[N][6];
	point_jacobian_t<N>(P.model, kd, frame.body, x, JT);
	if (out.M)
#pragma unroll
		for (int r = 0; r < N; r++)
#pragma unroll
			for (int c = 0; c < N; c++) out.M[(int64_t)(r * N + c) * NR + i] = kd.M[r][c];
	if (out.J)
#pragma unroll
		for (int r = 0; r < 6; r++)
#pragma unroll
			for (int c = 0; c < N; c++) out.J[(int64_t)(r * N + c) * NR + i] = JT[c][r];
	if (out.x)
#pragma unroll
		for (int k = 0; k < 3; k++) out.x[(int64_t)k * NR + i] = x[k];
	if (out.R)
#pragma unroll
		for (int k = 0; k < 9; k++) out.R[(int64_t)k * NR + i] = R[k];
	if (out.g)
#pragma unroll
		for (int j = 0; j < N; j++) out.g[(int64_t)j * NR + i] = kd.g[j];
}

#undef ST
}  // namespace osc
