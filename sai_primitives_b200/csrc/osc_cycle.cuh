// The fused control-cycle kernel: one robot per thread, FP64, no tensor cores.
//
//   kinematics/dynamics -> MotionForceTask model + torque -> JointTask model + torque
//   -> RobotController stacking (previous-torque feed-forward, saturation, gravity)
//
// Reference path being replaced (one launch == these calls for N robots):
//   SaiModel::updateModel                               examples/05-...cpp:143-145
//   RobotController::updateControllerTaskModels         src/RobotController.cpp:68-77
//     MotionForceTask::updateTaskModel                  src/tasks/MotionForceTask.cpp:247-268
//       SingularityHandler::updateTaskModel             src/tasks/SingularityHandler.cpp:75-228
//     JointTask::updateTaskModel                        src/tasks/JointTask.cpp:218-283
//   RobotController::computeControlTorques              src/RobotController.cpp:79-118
//     MotionForceTask::computeTorques                   src/tasks/MotionForceTask.cpp:270-509
//       SingularityHandler::computeTorques              src/tasks/SingularityHandler.cpp:297-309
//     JointTask::computeTorques                         src/tasks/JointTask.cpp:285-356
//
// Algebra (validated on the CPU by tests/test_algebra_proto.py):
// with M = L L^T every task Jacobian is whitened, X = L^-1 J^T, so that
// J M^-1 J^T = X^T X and the dynamically consistent null space becomes the orthogonal
// complement of range(X).  A Householder QR of X gives Lambda = (R^T R)^-1 and, through
// its reflectors, an orthonormal basis Q_perp of that complement; a full joint task in
// the null space is then  tau = L Q_perp G^-1 W^T(.)  with  W = L^-T Q_perp, G = W^T W.
// No singular value decomposition is needed on this (non-singular) branch: the branch
// decision uses the sound spectral test in osc_math.cuh; robots that fail it are flagged
// and handled by the SVD path.
#pragma once
#include "osc_kindyn.cuh"
#include "osc_launch.h"
#include "osc_tasks.cuh"

namespace osc {

// Optional re-convergence of the warps of a block at phase boundaries (so that they walk the large, fully
// unrolled instruction stream together).  Measured on B200: no gain -- the instruction-fetch stalls of this
// kernel are latency-, not bandwidth-bound (profiles/r01_v1_summary.md) -- so it is off; only legal when the
// grid covers the batch exactly (the kernel returns early for out-of-range threads).
#ifndef OSC_PHASE_SYNC
#define OSC_PHASE_SYNC 0
#endif
#if OSC_PHASE_SYNC
#define OSC_SYNC() __syncthreads()
#else
#define OSC_SYNC() ((void)0)
#endif

template <int N>
DEVI void bie_cholesky(const double (&M)[N][N], double thr, double (&Lb)[N][N]) {
#pragma unroll
	for (int r = 0; r < N; r++)
#pragma unroll
		for (int c = 0; c <= r; c++) Lb[r][c] = (r == c && M[r][c] < thr) ? thr : M[r][c];
	cholesky_lower<N>(Lb);
}

// Signature <N, R, HAS_JT>:  R = rank of a leading MotionForceTask (0: none),
// HAS_JT = a full JointTask closes the hierarchy.
#ifndef OSC_MIN_BLOCKS
#define OSC_MIN_BLOCKS 1
#endif
template <int N, int R, bool HAS_JT>
__global__ void __launch_bounds__(kCycleBlock, OSC_MIN_BLOCKS) osc_cycle_kernel(const __grid_constant__ OscProgram P) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t NR = P.n_robots;
	if (i == 0 && P.sing_count) P.sing_count[P.sing_parity ^ 1] = 0;
	if (i >= NR) return;
	const DevModel& mdl = P.model;

	double q[N], dq[N];
#pragma unroll
	for (int j = 0; j < N; j++) {
		q[j] = P.q[(int64_t)j * NR + i];
		dq[j] = P.dq[(int64_t)j * NR + i];
	}
	KinDyn<N> kd;
	forward_kinematics<N>(mdl, q, kd);
	OSC_SYNC();
	if (P.gravity_comp)
		mass_matrix<N, true>(mdl, kd);
	else
		mass_matrix<N, false>(mdl, kd);
	OSC_SYNC();

	double tau[N];
#pragma unroll
	for (int j = 0; j < N; j++) tau[j] = 0.0;
	uint32_t status = 0;

	// M = L L^T
	double L[N][N];
	if (R > 0) {
#pragma unroll
		for (int r = 0; r < N; r++)
#pragma unroll
			for (int c = 0; c <= r; c++) L[r][c] = kd.M[r][c];
		cholesky_lower<N>(L);
	}
	OSC_SYNC();

	// Householder data of the motion-force task (kept for the joint task's null space)
	double X[N][R > 0 ? R : 1];
	double vhead[R > 0 ? R : 1], beta[R > 0 ? R : 1];

	if constexpr (R > 0) {
		const DevMft& t = P.mft[0];
		const osc_mft_params& p = t.p;
		double x[3], Rc[9];
		frame_pose<N>(kd, t.body, t.ctrl_R, t.ctrl_t, x, Rc);
		double JT0[N][6];
		point_jacobian_t<N>(mdl, kd, t.body, x, JT0);
		// task rows J_t = B^T J0, stored transposed (N x R)
		double JtT[N][R];
#pragma unroll
		for (int j = 0; j < N; j++)
#pragma unroll
			for (int a = 0; a < R; a++) {
				if (t.full) {
					JtT[j][a] = JT0[j][a < 6 ? a : 0];
				} else {
					double s = 0.0;
#pragma unroll
					for (int k = 0; k < 6; k++) s += JT0[j][k] * t.B[k][a];
					JtT[j][a] = s;
				}
			}
		// Branch decision of SingularityHandler::updateTaskModel (:83-105), taken before any task state is touched:
		// robots that are not provably non-singular are appended (warp-aggregated) to the list of the SVD kernel
		// and leave this kernel.
		const bool flagged = !sound_nonsingular<N, R>(JtT, p.s_max, p.s_abs_tol);
		{
			const unsigned act = __activemask();
			const unsigned m = __ballot_sync(act, flagged);
			if (flagged) {
				const int lane = threadIdx.x & 31;
				const int leader = __ffs(m) - 1;
				int base = 0;
				if (lane == leader) base = atomicAdd(&P.sing_count[P.sing_parity], __popc(m));
				base = __shfl_sync(m, base, leader);
				P.sing_list[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)i;
				return;
			}
		}
		// classifySingularity with an empty singular range clears the handler memory (:239-245); only robots that
		// were singular at the previous update have anything to clear
		if (P.update_models && t.ist[(int64_t)MI_N_TYPES * NR + i] != 0) {
			t.ist[(int64_t)MI_N_TYPES * NR + i] = 0;
			t.ist[(int64_t)MI_T1_COUNTER * NR + i] = 0;
			t.ist[(int64_t)MI_T2_COUNTER * NR + i] = 0;
			t.ist[(int64_t)MI_HIST_HEAD * NR + i] = 0;
			t.ist[(int64_t)MI_HIST_SIZE * NR + i] = 0;
		}
	OSC_SYNC();

		double fstar[6], F[6];
		mft_control_law<N>(t, NR, i, x, Rc, JT0, dq, P.write_observers != 0, fstar, F, status);
	OSC_SYNC();
		double yf[R], yF[R];
#pragma unroll
		for (int a = 0; a < R; a++) {
			if (t.full) {
				yf[a] = fstar[a < 6 ? a : 0];
				yF[a] = F[a < 6 ? a : 0];
			} else {
				double s1 = 0.0, s2 = 0.0;
#pragma unroll
				for (int k = 0; k < 6; k++) {
					s1 += t.B[k][a] * fstar[k];
					s2 += t.B[k][a] * F[k];
				}
				yf[a] = s1;
				yF[a] = s2;
			}
		}
		const bool need_qr = HAS_JT || (p.dynamic_decoupling_type == OSC_FULL_DYNAMIC_DECOUPLING);
		if (need_qr) {
#pragma unroll
			for (int a = 0; a < R; a++) {
				double col[N];
#pragma unroll
				for (int j = 0; j < N; j++) col[j] = JtT[j][a];
				solve_lower<N>(L, col);
#pragma unroll
				for (int j = 0; j < N; j++) X[j][a] = col[j];
			}
			householder_qr<N, R, 0>(X, vhead, beta);
		}
	OSC_SYNC();
		if (p.dynamic_decoupling_type == OSC_FULL_DYNAMIC_DECOUPLING) {
			solve_rtr<N, R, 0>(X, yf);
		} else if (p.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES) {
			double Lb[N][N];
			bie_cholesky<N>(kd.M, p.bie_threshold, Lb);
			double Wb[N][R];
	OSC_SYNC();
#pragma unroll
			for (int a = 0; a < R; a++) {
				double col[N];
#pragma unroll
				for (int j = 0; j < N; j++) col[j] = JtT[j][a];
				solve_lower<N>(Lb, col);
#pragma unroll
				for (int j = 0; j < N; j++) Wb[j][a] = col[j];
			}
			double Ab[R][R];
#pragma unroll
			for (int a = 0; a < R; a++)
#pragma unroll
				for (int b = 0; b <= a; b++) {
					double s = 0.0;
#pragma unroll
					for (int j = 0; j < N; j++) s += Wb[j][a] * Wb[j][b];
					Ab[a][b] = s;
				}
			cholesky_lower<R>(Ab);
	OSC_SYNC();
			solve_spd<R>(Ab, yf);
		}  // IMPEDANCE: Lambda_modified = I
#pragma unroll
		for (int j = 0; j < N; j++) {
			double s = 0.0;
#pragma unroll
			for (int a = 0; a < R; a++) s += JtT[j][a] * (yf[a] + yF[a]);
			tau[j] += s;
		}
	}

	if constexpr (HAS_JT) {
		const DevJt& t = P.jt[0];
		const osc_joint_params& p = t.p;
		constexpr int Mn = N - R;  // dimension of the remaining null space
		if constexpr (R == 0) {
			// first task, N_prec = I, range = I:  tau = M qdd_d + M_mod t   (JointTask.cpp:348-355)
			double pid[N], acc[N];
			joint_control_law<N, N>(t, NR, i, q, dq, pid, acc);
#pragma unroll
			for (int r = 0; r < N; r++) {
				double s = 0.0;
#pragma unroll
				for (int c = 0; c < N; c++) {
					double mod;
					if (p.dynamic_decoupling_type == OSC_FULL_DYNAMIC_DECOUPLING)
						mod = kd.M[r][c];
					else if (p.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES)
						mod = (r == c && kd.M[r][c] < p.bie_threshold) ? p.bie_threshold : kd.M[r][c];
					else
						mod = (r == c) ? 1.0 : 0.0;
					s += kd.M[r][c] * acc[c] + mod * pid[c];
				}
				tau[r] += s;
			}
		} else if constexpr (Mn == 0) {
			status |= OSC_STATUS_ZERO_RANGE;  // JointTask.cpp:234-239, 302-306
		} else {
			double pid[N], acc[N];
			joint_control_law<N, N>(t, NR, i, q, dq, pid, acc);
			OSC_SYNC();
			// Q_perp = H_1..H_R [0; I],  W = L^-T Q_perp,  K = L Q_perp
			double W[N][Mn], K[N][Mn];
#pragma unroll
			for (int a = 0; a < Mn; a++) {
				double e[N];
#pragma unroll
				for (int j = 0; j < N; j++) e[j] = (j == R + a) ? 1.0 : 0.0;
				apply_q<N, R, 0>(X, vhead, beta, e);
				double k[N];
				mul_lower<N>(L, e, k);
				solve_lower_t<N>(L, e);
#pragma unroll
				for (int j = 0; j < N; j++) {
					W[j][a] = e[j];
					K[j][a] = k[j];
				}
			}
			double G[Mn][Mn];
			double nrm_chk = 0.0;
#pragma unroll
			for (int a = 0; a < Mn; a++)
#pragma unroll
				for (int b = 0; b <= a; b++) {
					double s = 0.0, kk = 0.0;
#pragma unroll
					for (int j = 0; j < N; j++) {
						s += W[j][a] * W[j][b];
						kk += K[j][a] * K[j][b];
					}
					G[a][b] = s;
					nrm_chk += (a == b) ? s * kk : 2.0 * s * kk;
				}
			// ||N_prec||_F^2 = tr(G K^T K) must stay < 1e6 for the reference's 1e-3 range tolerance
			if (!(nrm_chk < 1.0e6)) status |= OSC_STATUS_UNHANDLED;
			cholesky_lower<Mn>(G);
	OSC_SYNC();
			// u = W^T (qdd_d - M^-1 tau_prec)
			double rhs[N];
#pragma unroll
			for (int j = 0; j < N; j++) rhs[j] = tau[j];
			if (P.use_prev_torques) {
				solve_spd<N>(L, rhs);
#pragma unroll
				for (int j = 0; j < N; j++) rhs[j] = acc[j] - rhs[j];
			} else {
#pragma unroll
				for (int j = 0; j < N; j++) rhs[j] = acc[j];
			}
			double z1[Mn], z2[Mn];
#pragma unroll
			for (int a = 0; a < Mn; a++) {
				double s1 = 0.0, s2 = 0.0;
#pragma unroll
				for (int j = 0; j < N; j++) {
					s1 += W[j][a] * rhs[j];
					s2 += W[j][a] * pid[j];
				}
				z1[a] = s1;
				z2[a] = s2;
			}
			solve_spd<Mn>(G, z1);
			if (p.dynamic_decoupling_type == OSC_FULL_DYNAMIC_DECOUPLING) {
				solve_spd<Mn>(G, z2);
			} else if (p.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES) {
				solve_spd<Mn>(G, z2);
				double Lb[N][N];
				bie_cholesky<N>(kd.M, p.bie_threshold, Lb);
				double Z[N][Mn];
	OSC_SYNC();
#pragma unroll
				for (int a = 0; a < Mn; a++) {
					double col[N];
#pragma unroll
					for (int j = 0; j < N; j++) col[j] = K[j][a];
					solve_lower<N>(Lb, col);
#pragma unroll
					for (int j = 0; j < N; j++) Z[j][a] = col[j];
				}
				double H[Mn][Mn];
#pragma unroll
				for (int a = 0; a < Mn; a++)
#pragma unroll
					for (int b = 0; b <= a; b++) {
						double s = 0.0;
#pragma unroll
						for (int j = 0; j < N; j++) s += Z[j][a] * Z[j][b];
						H[a][b] = s;
					}
				cholesky_lower<Mn>(H);
				solve_spd<Mn>(H, z2);
			}  // IMPEDANCE: z2 = W^T t
#pragma unroll
			for (int j = 0; j < N; j++) {
				double s = 0.0;
#pragma unroll
				for (int a = 0; a < Mn; a++) s += K[j][a] * (z1[a] + z2[a]);
				tau[j] += s;
			}
		}
	}

	// RobotController::computeControlTorques tail (RobotController.cpp:86-116)
	if (P.torque_saturation) {
#pragma unroll
		for (int j = 0; j < N; j++) {
			if (tau[j] > mdl.effort[j])
				tau[j] = mdl.effort[j];
			else if (tau[j] < -mdl.effort[j])
				tau[j] = -mdl.effort[j];
		}
	}
	if (P.gravity_comp) {
#pragma unroll
		for (int j = 0; j < N; j++) tau[j] += kd.g[j];
	}
	if (status & OSC_STATUS_UNHANDLED) {
#pragma unroll
		for (int j = 0; j < N; j++) tau[j] = __longlong_as_double(0x7ff8000000000000LL);
	}
#pragma unroll
	for (int j = 0; j < N; j++) P.tau[(int64_t)j * NR + i] = tau[j];
	P.status[i] = status;
}

}  // namespace osc
