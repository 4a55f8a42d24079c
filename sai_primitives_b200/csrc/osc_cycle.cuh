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
// The bounded-inertia variant M_BIE = M + diag(d) is a rank-one update M + h h^T whenever a single
// diagonal entry is clamped (always the case for the Panda: SURVEY.md Appendix D), handled by
// Sherman-Morrison on the factors already at hand; several clamped entries take the general route.
// No singular value decomposition is needed on this (non-singular) branch: the branch
// decision uses the sound spectral test in osc_math.cuh; robots that fail it are handed to the
// general-path kernel (osc_singular.cuh) through a compacted list.
#pragma once
#include "osc_kindyn.cuh"
#include "osc_launch.h"
#include "osc_tasks.cuh"

namespace osc {

#ifndef OSC_MIN_BLOCKS
#define OSC_MIN_BLOCKS 1
#endif

// h_i = sqrt(max(thr - M_ii, 0)):  M_BIE = M + diag(h_i^2).  Returns the number of clamped entries.
template <int N>
DEVI int bie_shift(const double (&Mdiag)[N], double thr, double (&h)[N]) {
	int k = 0;
#pragma unroll
	for (int j = 0; j < N; j++) {
		const double d = thr - Mdiag[j];
		h[j] = (d > 0.0) ? sqrt(d) : 0.0;
		k += (d > 0.0) ? 1 : 0;
	}
	return k;
}

// Cholesky of M_BIE re-assembled from M = L L^T (the lower triangle of M is overwritten by its factor):
// general route for two or more clamped entries.
template <int N>
DEVI void bie_cholesky_from_factor(const double (&L)[N][N], const double (&h)[N], double (&Lb)[N][N], double (&invdb)[N]) {
#pragma unroll
	for (int r = 0; r < N; r++)
#pragma unroll
		for (int c = 0; c <= r; c++) {
			double s = 0.0;
#pragma unroll
			for (int k = 0; k <= c; k++) s += L[r][k] * L[c][k];
			Lb[r][c] = (r == c) ? s + h[r] * h[r] : s;
		}
	cholesky_lower<N>(Lb, invdb);
}

// Signature <N, R, HAS_JT, FULL>:  R = rank of a leading MotionForceTask (0: none), FULL = that task controls all six
// directions (B = I), HAS_JT = a full JointTask closes the hierarchy.
// Dynamic shared memory: 9 N doubles per thread (body orientations between the two kinematics passes),
// element e of thread t at  sm[e * blockDim.x + t]  (conflict-free, 8-byte interleave).
template <int N, int R, bool HAS_JT, bool FULL>
__global__ void __launch_bounds__(kCycleBlock, OSC_MIN_BLOCKS) osc_cycle_kernel(const __grid_constant__ OscProgram P) {
	extern __shared__ double sm[];
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t NR = P.n_robots;
	if (i == 0 && P.sing_count) P.sing_count[P.sing_parity ^ 1] = 0;
	if (i >= NR) return;
	const DevModel& mdl = P.model;
	double* smt = sm + threadIdx.x;
	const int sms = blockDim.x;

	double q[N], dq[N];
#pragma unroll
	for (int j = 0; j < N; j++) {
		q[j] = P.q[(int64_t)j * NR + i];
		dq[j] = P.dq[(int64_t)j * NR + i];
	}
	KinDynS<N> kd;
	forward_kinematics_s<N>(mdl, q, kd, smt, sms);
	if (P.gravity_comp)
		mass_matrix_s<N, true>(mdl, kd, smt, sms);
	else
		mass_matrix_s<N, false>(mdl, kd, smt, sms);

	double tau[N];
#pragma unroll
	for (int j = 0; j < N; j++) tau[j] = 0.0;
	uint32_t status = 0;

	if constexpr (R == 0) {
		// ---- a full JointTask alone: N_prec = I, range = I:  tau = M qdd_d + M_mod t   (JointTask.cpp:348-355)
		const DevJt& t = P.jt[0];
		const osc_joint_params& p = t.p;
		double pid[N], acc[N];
		joint_control_law<N, N>(t, NR, i, q, dq, pid, acc);
#pragma unroll
		for (int r = 0; r < N; r++) {
			double s = 0.0;
#pragma unroll
			for (int c = 0; c < N; c++) {
				const double m = (c <= r) ? kd.M[r][c] : kd.M[c][r];
				double mod;
				if (p.dynamic_decoupling_type == OSC_FULL_DYNAMIC_DECOUPLING)
					mod = m;
				else if (p.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES)
					mod = (r == c && m < p.bie_threshold) ? p.bie_threshold : m;
				else
					mod = (r == c) ? 1.0 : 0.0;
				s += m * acc[c] + mod * pid[c];
			}
			tau[r] += s;
		}
	} else {
		const DevMft& t = P.mft[0];
		const osc_mft_params& p = t.p;
		double x[3], Rc[9];
		frame_pose_s<N>(kd, t.body, t.ctrl_R, t.ctrl_t, x, Rc, smt, sms);
		// task rows J_t = B^T J0, stored transposed (N x R); for a full task J_t = J0
		double JtT[N][R];
		double v[3] = {0, 0, 0}, w[3] = {0, 0, 0};	// J0 dq
#pragma unroll
		for (int j = 0; j < N; j++) {
			double c6[6];
			jacobian_column<N>(mdl, kd, t.body, x, j, c6);
#pragma unroll
			for (int k = 0; k < 3; k++) {
				v[k] += c6[k] * dq[j];
				w[k] += c6[3 + k] * dq[j];
			}
			if constexpr (FULL) {
#pragma unroll
				for (int a = 0; a < R; a++) JtT[j][a] = c6[a < 6 ? a : 0];
			} else {
#pragma unroll
				for (int a = 0; a < R; a++) {
					double s = 0.0;
#pragma unroll
					for (int k = 0; k < 6; k++) s += c6[k] * t.B[k][a];
					JtT[j][a] = s;
				}
			}
		}
		// Branch decision of SingularityHandler::updateTaskModel (:83-105), taken before any task state is touched:
		// robots that are not provably non-singular are appended (warp-aggregated) to the list of the general-path
		// kernel and leave this kernel.
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

		// M = L L^T in place; the diagonal of M is all the bounded-inertia variant needs later
		double Mdiag[N], invd[N];
#pragma unroll
		for (int j = 0; j < N; j++) Mdiag[j] = kd.M[j][j];
		cholesky_lower<N>(kd.M, invd);
		double(&L)[N][N] = kd.M;

		double fstar[6], F[6];
		const bool has_F = mft_control_law(t, NR, i, x, Rc, v, w, P.write_observers != 0, fstar, F, status);
		double yf[R], yF[R];
#pragma unroll
		for (int a = 0; a < R; a++) {
			if constexpr (FULL) {
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

		// X = L^-1 J_t^T, all R columns row by row
		double X[N][R];
#pragma unroll
		for (int r = 0; r < N; r++)
#pragma unroll
			for (int a = 0; a < R; a++) {
				double s = JtT[r][a];
#pragma unroll
				for (int k = 0; k < r; k++) s -= L[r][k] * X[k][a];
				X[r][a] = s * invd[r];
			}
		// bounded inertia estimates: M_BIE = M + diag(h^2); with one clamped entry: M + h h^T, g = L^-1 h, z = J M^-1 h = X^T g
		const int dec = p.dynamic_decoupling_type;
		int kclamp = 0;
		double z[R], mu = 0.0;
		double h[N];
		if (dec == OSC_BOUNDED_INERTIA_ESTIMATES) {
			kclamp = bie_shift<N>(Mdiag, p.bie_threshold, h);
			if (kclamp == 1) {
				double g[N];
#pragma unroll
				for (int j = 0; j < N; j++) g[j] = h[j];
				solve_lower<N>(L, invd, g);
#pragma unroll
				for (int j = 0; j < N; j++) mu += g[j] * g[j];
#pragma unroll
				for (int a = 0; a < R; a++) {
					double s = 0.0;
#pragma unroll
					for (int j = 0; j < N; j++) s += X[j][a] * g[j];
					z[a] = s;
				}
			}
		}
		double vhead[R], beta[R], rinv[R];
		householder_qr<N, R, 0>(X, vhead, beta, rinv);

		if (dec == OSC_FULL_DYNAMIC_DECOUPLING || (dec == OSC_BOUNDED_INERTIA_ESTIMATES && kclamp == 0)) {
			solve_rtr<N, R, 0>(X, rinv, yf);
		} else if (dec == OSC_BOUNDED_INERTIA_ESTIMATES && kclamp == 1) {
			// (A - z z^T / (1 + mu))^-1 y = A^-1 y + (A^-1 z) (z . A^-1 y) / ((1 + mu) - z . A^-1 z)
			double sz[R];
#pragma unroll
			for (int a = 0; a < R; a++) sz[a] = z[a];
			solve_rtr<N, R, 0>(X, rinv, sz);
			solve_rtr<N, R, 0>(X, rinv, yf);
			double zs = 0.0, zt = 0.0;
#pragma unroll
			for (int a = 0; a < R; a++) {
				zs += z[a] * sz[a];
				zt += z[a] * yf[a];
			}
			const double f = zt / ((1.0 + mu) - zs);
#pragma unroll
			for (int a = 0; a < R; a++) yf[a] += sz[a] * f;
		} else if (dec == OSC_BOUNDED_INERTIA_ESTIMATES) {
			// general route: A_b = (L_b^-1 J^T)^T (L_b^-1 J^T)
			double Lb[N][N], invdb[N];
			bie_cholesky_from_factor<N>(L, h, Lb, invdb);
			double Wb[N][R];
#pragma unroll
			for (int r = 0; r < N; r++)
#pragma unroll
				for (int a = 0; a < R; a++) {
					double s = JtT[r][a];
#pragma unroll
					for (int k = 0; k < r; k++) s -= Lb[r][k] * Wb[k][a];
					Wb[r][a] = s * invdb[r];
				}
			double Ab[R][R], invda[R];
#pragma unroll
			for (int a = 0; a < R; a++)
#pragma unroll
				for (int b = 0; b <= a; b++) {
					double s = 0.0;
#pragma unroll
					for (int j = 0; j < N; j++) s += Wb[j][a] * Wb[j][b];
					Ab[a][b] = s;
				}
			cholesky_lower<R>(Ab, invda);
			solve_spd<R>(Ab, invda, yf);
		}  // IMPEDANCE: Lambda_modified = I
#pragma unroll
		for (int j = 0; j < N; j++) {
			double s = 0.0;
#pragma unroll
			for (int a = 0; a < R; a++) s += JtT[j][a] * (has_F ? yf[a] + yF[a] : yf[a]);
			tau[j] += s;
		}

		if constexpr (HAS_JT) {
			const DevJt& jt = P.jt[0];
			const osc_joint_params& jp = jt.p;
			constexpr int Mn = N - R;  // dimension of the remaining null space
			if constexpr (Mn == 0) {
				status |= OSC_STATUS_ZERO_RANGE;  // JointTask.cpp:234-239, 302-306
			} else {
				double pid[N], acc[N];
				joint_control_law<N, N>(jt, NR, i, q, dq, pid, acc);
				// Q_perp = H_1..H_R [0; I],  W = L^-T Q_perp,  K = L Q_perp
				double Qp[N][Mn], W[N][Mn], K[N][Mn];
#pragma unroll
				for (int a = 0; a < Mn; a++) {
					double e[N];
#pragma unroll
					for (int j = 0; j < N; j++) e[j] = (j == R + a) ? 1.0 : 0.0;
					apply_q<N, R, 0>(X, vhead, beta, e);
					double k[N];
					mul_lower<N>(L, e, k);
#pragma unroll
					for (int j = 0; j < N; j++) Qp[j][a] = e[j];
					solve_lower_t<N>(L, invd, e);
#pragma unroll
					for (int j = 0; j < N; j++) {
						W[j][a] = e[j];
						K[j][a] = k[j];
					}
				}
				double G[Mn][Mn], invg[Mn];
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
				cholesky_lower<Mn>(G, invg);
				// u = W^T (qdd_d - M^-1 tau_prec)
				double rhs[N];
				if (P.use_prev_torques) {
#pragma unroll
					for (int j = 0; j < N; j++) rhs[j] = tau[j];
					solve_spd<N>(L, invd, rhs);
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
				solve_spd<Mn>(G, invg, z1);
				const int jdec = jp.dynamic_decoupling_type;
				if (jdec == OSC_FULL_DYNAMIC_DECOUPLING) {
					solve_spd<Mn>(G, invg, z2);
				} else if (jdec == OSC_BOUNDED_INERTIA_ESTIMATES) {
					solve_spd<Mn>(G, invg, z2);
					double hj[N];
					const int kj = bie_shift<N>(Mdiag, jp.bie_threshold, hj);
					if (kj == 1) {
						// H = K^T M_b^-1 K = I - c c^T / (1 + mu),  c = Q_perp^T L^-1 h;  H^-1 z = z + c (c . z) / ((1 + mu) - c . c)
						double g[N];
#pragma unroll
						for (int j = 0; j < N; j++) g[j] = hj[j];
						solve_lower<N>(L, invd, g);
						double muj = 0.0;
#pragma unroll
						for (int j = 0; j < N; j++) muj += g[j] * g[j];
						double c[Mn], cc = 0.0, cz = 0.0;
#pragma unroll
						for (int a = 0; a < Mn; a++) {
							double s = 0.0;
#pragma unroll
							for (int j = 0; j < N; j++) s += Qp[j][a] * g[j];
							c[a] = s;
							cc += s * s;
							cz += s * z2[a];
						}
						const double f = cz / ((1.0 + muj) - cc);
#pragma unroll
						for (int a = 0; a < Mn; a++) z2[a] += c[a] * f;
					} else if (kj >= 2) {
						double Lb[N][N], invdb[N];
						bie_cholesky_from_factor<N>(L, hj, Lb, invdb);
						double Z[N][Mn];
#pragma unroll
						for (int a = 0; a < Mn; a++) {
							double col[N];
#pragma unroll
							for (int j = 0; j < N; j++) col[j] = K[j][a];
							solve_lower<N>(Lb, invdb, col);
#pragma unroll
							for (int j = 0; j < N; j++) Z[j][a] = col[j];
						}
						double H[Mn][Mn], invh[Mn];
#pragma unroll
						for (int a = 0; a < Mn; a++)
#pragma unroll
							for (int b = 0; b <= a; b++) {
								double s = 0.0;
#pragma unroll
								for (int j = 0; j < N; j++) s += Z[j][a] * Z[j][b];
								H[a][b] = s;
							}
						cholesky_lower<Mn>(H, invh);
						solve_spd<Mn>(H, invh, z2);
					}
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
