// Unrolled singular / blending path for the flagship hierarchy [full 6-DoF MotionForceTask first, optional full
// JointTask]: the robots the fast kernel hands over are re-evaluated here one per thread with everything in
// registers, instead of the rolled general-path code (osc_singular.cuh), which stays the fallback for what this
// file does not specialise (three or more singular directions, a fully singular task, several clamped inertia
// entries, singularity handling switched off).
//
// Reference: SingularityHandler::updateTaskModel / classifySingularity / computeTorques
// (src/tasks/SingularityHandler.cpp:75-368) and JointTask with the resulting null space (src/tasks/JointTask.cpp:218-356).
//
// Algebra (same whitening as the fast path).  With M = L L^T, X = L^-1 J^T (n x 6) and the
// eigen-decomposition J J^T = U S^2 U^T (cyclic Jacobi on the 6 x 6 Gram matrix):
//   U = [U_ns | U_s] (first NNS / last NS columns),  V_s = J^T U_s S_s^-1 = L X U_s S_s^-1
//   X U = Q R' (Householder):  J_ns M^-1 J_ns^T = R'_11^T R'_11,   J_s M^-1 J_s^T = R'_12^T R'_12 + R'_22^T R'_22,
//   posture Jacobian J_post = V_s^T N_ns:  L^-1 J_post^T = Q [0; R'_22; 0] S_s^-1,  so
//   J_post M^-1 J_post^T = S_s^-1 R'_22^T R'_22 S_s^-1,  and N = N_js N_ns is L^-T (I - Q_6 Q_6^T) L^T:
//   the joint task sees exactly the complement Q e_7 of the fast path.
#pragma once
#include "osc_kindyn.cuh"
#include "osc_singular.cuh"
#include "osc_tasks.cuh"

namespace osc {

// Jacobi eigen-decomposition of a symmetric 6 x 6 matrix held in registers: G -> diag(lambda), U <- eigenvectors (columns).
// Parallel (round-robin) ordering: a sweep is five rounds of three rotations on disjoint index pairs.  The three rotation
// angles of a round come from the same matrix, so their scalar chains (reciprocal, two square roots: ~30 dependent FP64
// operations each) overlap instead of following one another, and the rotations are branch-free (c = 1, s = 0 where the
// off-diagonal entry is already negligible): no divergence inside a sweep.  The sweep loop is per thread (warp
// divergence = the slowest lane).
template <int P, int Q>
DEVI void jacobi_angle(const double (&G)[6][6], double& c, double& s) {
	const double gpq = G[P][Q];
	const bool rot = fabs(gpq) > 1e-18 * (fabs(G[P][P]) + fabs(G[Q][Q]));
	const double g = rot ? gpq : 1.0;
	// lean reciprocal / square roots (osc_math.cuh): g is non-zero and both radicands are >= 1
	const double theta = (G[Q][Q] - G[P][P]) * (0.5 * rcp_nz(g));
	const double t = (theta >= 0.0 ? 1.0 : -1.0) * rcp_nz(fabs(theta) + sqrt_pos(theta * theta + 1.0));
	const double cc = rsqrt_pos(t * t + 1.0);
	c = rot ? cc : 1.0;
	s = rot ? t * cc : 0.0;
}
template <int P, int Q>
DEVI void jacobi_cols(double (&A)[6][6], double c, double s) {	// A <- A J(P, Q)
#pragma unroll
	for (int k = 0; k < 6; k++) {
		const double akp = A[k][P], akq = A[k][Q];
		A[k][P] = c * akp - s * akq;
		A[k][Q] = s * akp + c * akq;
	}
}
template <int P, int Q>
DEVI void jacobi_rows(double (&A)[6][6], double c, double s) {	// A <- J(P, Q)^T A
#pragma unroll
	for (int k = 0; k < 6; k++) {
		const double apk = A[P][k], aqk = A[Q][k];
		A[P][k] = c * apk - s * aqk;
		A[Q][k] = s * apk + c * aqk;
	}
}
template <int P0, int Q0, int P1, int Q1, int P2, int Q2>
DEVI void jacobi_round(double (&G)[6][6], double (&U)[6][6]) {
	double c0, s0, c1, s1, c2, s2;
	jacobi_angle<P0, Q0>(G, c0, s0);
	jacobi_angle<P1, Q1>(G, c1, s1);
	jacobi_angle<P2, Q2>(G, c2, s2);
	jacobi_cols<P0, Q0>(G, c0, s0);
	jacobi_cols<P1, Q1>(G, c1, s1);
	jacobi_cols<P2, Q2>(G, c2, s2);
	jacobi_rows<P0, Q0>(G, c0, s0);
	jacobi_rows<P1, Q1>(G, c1, s1);
	jacobi_rows<P2, Q2>(G, c2, s2);
	jacobi_cols<P0, Q0>(U, c0, s0);
	jacobi_cols<P1, Q1>(U, c1, s1);
	jacobi_cols<P2, Q2>(U, c2, s2);
}
DEVI void jacobi_eig6(double (&G)[6][6], double (&U)[6][6]) {
#pragma unroll
	for (int a = 0; a < 6; a++)
#pragma unroll
		for (int b = 0; b < 6; b++) U[a][b] = (a == b) ? 1.0 : 0.0;
	for (int sweep = 0; sweep < 30; sweep++) {
		double off = 0.0, dia = 0.0;
#pragma unroll
		for (int a = 0; a < 6; a++) {
			dia += G[a][a] * G[a][a];
#pragma unroll
			for (int b = 0; b < a; b++) off += G[a][b] * G[a][b];
		}
		// off-diagonal entries at 1e-15 of the diagonal scale: the rounding floor of the rotations is ~1e-16
		if (off <= 1e-30 * dia) break;
		jacobi_round<0, 5, 1, 4, 2, 3>(G, U);
		jacobi_round<0, 4, 3, 5, 1, 2>(G, U);
		jacobi_round<0, 3, 2, 4, 1, 5>(G, U);
		jacobi_round<0, 2, 1, 3, 4, 5>(G, U);
		jacobi_round<0, 1, 2, 5, 3, 4>(G, U);
	}
}

// (A - delta z z^T / kappa)^-1 y for A = Rb^T Rb given by its upper-triangular factor stored in T[C0 + i][C0 + j] (i <= j),
// size S; rinv = reciprocal diagonal.  use_sm = false gives plain A^-1 y.
template <int N, int S, int C0>
DEVI void solve_block_sm(const double (&T)[N][6], const double (&rinv)[6], double (&y)[S], const double (&z)[S], bool use_sm, double delta,
						 double kappa) {
	auto solve = [&](double(&x)[S]) {
#pragma unroll
		for (int i = 0; i < S; i++) {
			double s = x[i];
#pragma unroll
			for (int k = 0; k < i; k++) s -= T[C0 + k][C0 + i] * x[k];
			x[i] = s * rinv[C0 + i];
		}
#pragma unroll
		for (int i = S - 1; i >= 0; i--) {
			double s = x[i];
#pragma unroll
			for (int k = i + 1; k < S; k++) s -= T[C0 + i][C0 + k] * x[k];
			x[i] = s * rinv[C0 + i];
		}
	};
	solve(y);
	if (use_sm) {
		double sz[S];
#pragma unroll
		for (int a = 0; a < S; a++) sz[a] = z[a];
		solve(sz);
		double zs = 0.0, zt = 0.0;
#pragma unroll
		for (int a = 0; a < S; a++) {
			zs += z[a] * sz[a];
			zt += z[a] * y[a];
		}
		const double f = delta * zt / (kappa - delta * zs);
#pragma unroll
		for (int a = 0; a < S; a++) y[a] += sz[a] * f;
	}
}

// Everything after the eigen-decomposition, for NS singular directions (NS = 0: the thin band of non-singular robots
// the sound test of the fast kernel rejects).  Returns false when the case must go to the general path.
template <int N, int NS, bool HAS_JT>
DEVI bool blend_path(const OscProgram& P, int64_t i, const double (&q)[N], const double (&dq)[N], KinDyn<N>& kd, const double (&L)[N][N],
					 const double (&invd)[N], const double (&Mdiag)[N], const double (&JT0)[N][6], const double (&Xw)[N][6],
					 const double (&U)[6][6], const double (&sig)[6], double alpha, const double x[3], const double Rc[9], uint32_t status) {
	constexpr int NNS = 6 - NS;
	const int64_t NR = P.n_robots;
	const DevModel& mdl = P.model;
	const DevMft& t = P.mft[0];
	const osc_mft_params& p = t.p;
	double* st = t.st;
	int32_t* ist = t.ist;
	const int dec = p.dynamic_decoupling_type;

	// bounded inertia: one clamped entry -> rank-one (Sherman-Morrison); several -> general path
	double dclamp[N], delta = 0.0, mu = 0.0, g[N];
	int kclamp = 0;
	if (dec == OSC_BOUNDED_INERTIA_ESTIMATES) {
#pragma unroll
		for (int j = 0; j < N; j++) {
			const double dj = p.bie_threshold - Mdiag[j];
			dclamp[j] = (dj > 0.0) ? dj : 0.0;
			delta += dclamp[j];
			kclamp += (dj > 0.0) ? 1 : 0;
		}
		if (kclamp >= 2) return false;
	}
	const bool sm = (dec == OSC_BOUNDED_INERTIA_ESTIMATES) && kclamp == 1;
	if (sm) {
#pragma unroll
		for (int j = 0; j < N; j++) g[j] = (dclamp[j] > 0.0) ? 1.0 : 0.0;
		solve_lower<N>(L, invd, g);
#pragma unroll
		for (int j = 0; j < N; j++) mu += g[j] * g[j];
	}
	const double kappa = 1.0 + delta * mu;

	// right singular vectors of the singular directions: V_s = J^T U_s / sigma, oriented so that the largest-magnitude
	// entry is positive (sign convention of this repo); the matching U_s column follows the sign
	double Us[6][NS > 0 ? NS : 1], Vs[N][NS > 0 ? NS : 1];
	double Ue[6][6];  // [U_ns | U_s] with the final signs
#pragma unroll
	for (int a = 0; a < 6; a++)
#pragma unroll
		for (int c = 0; c < 6; c++) Ue[a][c] = U[a][c];
	if constexpr (NS > 0) {
#pragma unroll
		for (int c = 0; c < NS; c++) {
			double vmax = 0.0, amax = -1.0;
			const double inv = 1.0 / sig[NNS + c];
#pragma unroll
			for (int j = 0; j < N; j++) {
				double s = 0.0;
#pragma unroll
				for (int a = 0; a < 6; a++) s += JT0[j][a] * U[a][NNS + c];
				s *= inv;
				Vs[j][c] = s;
				if (fabs(s) > amax) {
					amax = fabs(s);
					vmax = s;
				}
			}
			const double sgn = (vmax < 0.0) ? -1.0 : 1.0;
#pragma unroll
			for (int j = 0; j < N; j++) Vs[j][c] *= sgn;
#pragma unroll
			for (int a = 0; a < 6; a++) {
				Us[a][c] = sgn * U[a][NNS + c];
				Ue[a][NNS + c] = Us[a][c];
			}
		}
	}

	// task velocity J0 dq (the Jacobian is not needed after this point)
	double v[3] = {0, 0, 0}, w[3] = {0, 0, 0};
#pragma unroll
	for (int j = 0; j < N; j++)
#pragma unroll
		for (int k = 0; k < 3; k++) {
			v[k] += JT0[j][k] * dq[j];
			w[k] += JT0[j][3 + k] * dq[j];
		}
	// ---- T = X U (its Householder QR follows below; X is not needed after this point)
	double T[N][6];
#pragma unroll
	for (int r = 0; r < N; r++)
#pragma unroll
		for (int c = 0; c < 6; c++) {
			double s = 0.0;
#pragma unroll
			for (int a = 0; a < 6; a++) s += Xw[r][a] * Ue[a][c];
			T[r][c] = s;
		}
	// z = (X U)^T g for the bounded-inertia rank-one update, before T is overwritten
	double zu[6];
#pragma unroll
	for (int c = 0; c < 6; c++) {
		double s = 0.0;
		if (sm) {
#pragma unroll
			for (int r = 0; r < N; r++) s += T[r][c] * g[r];
		}
		zu[c] = s;
	}

	// ---- classifySingularity (:230-295): memory of the handler
	int32_t c1 = ist[(int64_t)MI_T1_COUNTER * NR + i], c2 = ist[(int64_t)MI_T2_COUNTER * NR + i];
	const int32_t n_types_prev = ist[(int64_t)MI_N_TYPES * NR + i];
	int32_t hist_head = ist[(int64_t)MI_HIST_HEAD * NR + i], hist_size = ist[(int64_t)MI_HIST_SIZE * NR + i];
	double q_prior[N];
	const bool upd = P.update_models != 0;
	if (upd && (n_types_prev == 0 || c2 > c1)) {
#pragma unroll
		for (int j = 0; j < N; j++) {
			q_prior[j] = q[j];
			st[(int64_t)(MC_Q_PRIOR + j) * NR + i] = q[j];
			st[(int64_t)(MC_DQ_PRIOR + j) * NR + i] = dq[j];
		}
	} else {
#pragma unroll
		for (int j = 0; j < N; j++) q_prior[j] = st[(int64_t)(MC_Q_PRIOR + j) * NR + i];
	}
	int n_types = NS;
	if constexpr (NS == 0) {
		if (upd) {
			c1 = c2 = 0;
			hist_head = hist_size = 0;
		}
	} else {
		bool any_type1 = false;
#pragma unroll
		for (int c = 0; c < NS; c++) {
			double qq[N];
#pragma unroll
			for (int j = 0; j < N; j++) qq[j] = q[j] + p.perturb_step_size * Vs[j][c];
			double xp[3], Rp[9], dphi[3];
			pose_only<N>(mdl, qq, t.body, t.ctrl_R, t.ctrl_t, xp, Rp);
			orientation_error(Rp, Rc, dphi);
			double mot = 0.0;
#pragma unroll
			for (int k = 0; k < 3; k++) mot += (xp[k] - x[k]) * Us[k][c] + dphi[k] * Us[3 + k][c];
			if (fabs(mot) > p.type_1_tol) any_type1 = true;
		}
		if (upd) {
			const int pos = (hist_head + hist_size) % OSC_HIST_MAX;
			uint32_t* word = reinterpret_cast<uint32_t*>(&ist[(int64_t)(MI_HIST_BITS + pos / 32) * NR + i]);
			if (any_type1) {
				*word |= (1u << (pos % 32));
				c1++;
			} else {
				*word &= ~(1u << (pos % 32));
				c2++;
			}
			hist_size++;
			if (hist_size > p.buffer_size) {
				const uint32_t w0 = (uint32_t)ist[(int64_t)(MI_HIST_BITS + hist_head / 32) * NR + i];
				if ((w0 >> (hist_head % 32)) & 1u)
					c1--;
				else
					c2--;
				hist_head = (hist_head + 1) % OSC_HIST_MAX;
				hist_size--;
			}
		}
	}
	if (upd) {
		ist[(int64_t)MI_T1_COUNTER * NR + i] = c1;
		ist[(int64_t)MI_T2_COUNTER * NR + i] = c2;
		ist[(int64_t)MI_N_TYPES * NR + i] = n_types;
		ist[(int64_t)MI_HIST_HEAD * NR + i] = hist_head;
		ist[(int64_t)MI_HIST_SIZE * NR + i] = hist_size;
	} else {
		n_types = n_types_prev;
	}
	if constexpr (NS > 0) status |= OSC_STATUS_SINGULAR_PATH;

	// ---- control law (state update happens exactly once, here)
	double fstar[6], F[6];
	mft_control_law(t, NR, i, x, Rc, v, w, P.write_observers != 0, fstar, F, status);

	double vhead[6], beta[6], rinv[6];
	householder_qr<N, 6, 0>(T, vhead, beta, rinv);

	// coordinates of f*, F in the rotated task basis
	double au[6], bu[6];
#pragma unroll
	for (int c = 0; c < 6; c++) {
		double s1 = 0.0, s2 = 0.0;
#pragma unroll
		for (int a = 0; a < 6; a++) {
			s1 += Ue[a][c] * fstar[a];
			s2 += Ue[a][c] * F[a];
		}
		au[c] = s1;
		bu[c] = s2;
	}
	const bool types_nonempty = (n_types != 0);
	const bool impedance_plain = types_nonempty && dec == OSC_IMPEDANCE;

	// tau_ns = J_ns^T (Lambda_ns,mod a + b) = L Q [R'_11 c; 0]
	double e[N];
	{
		double cns[NNS], zns[NNS];
#pragma unroll
		for (int a = 0; a < NNS; a++) {
			cns[a] = au[a];
			zns[a] = zu[a];
		}
		if (dec != OSC_IMPEDANCE) solve_block_sm<N, NNS, 0>(T, rinv, cns, zns, sm, delta, kappa);
		(void)impedance_plain;
#pragma unroll
		for (int a = 0; a < NNS; a++) cns[a] += bu[a];
#pragma unroll
		for (int r = 0; r < N; r++) {
			double s = 0.0;
			if (r < NNS) {
#pragma unroll
				for (int a = r; a < NNS; a++) s += T[r][a] * cns[a];
			}
			e[r] = s;
		}
	}
	double tau[N];
	{
		double qe[N];
#pragma unroll
		for (int r = 0; r < N; r++) qe[r] = e[r];
		apply_q<N, 6, 0>(T, vhead, beta, qe);
		mul_lower<N>(L, qe, tau);
	}
	const bool handling = p.singularity_handling_enabled != 0;
	if constexpr (NS > 0) {
		if (!handling) return false;  // N = N_ns only: general path
		const bool blended = types_nonempty && dec != OSC_IMPEDANCE;
		if (blended) {
			// ---- singular task torques: J_s^T (Lambda_s,mod U_s^T f* + U_s^T F),  J_s M^-1 J_s^T = R'_12^T R'_12 + R'_22^T R'_22
			double As[NS][NS], ys[NS];
#pragma unroll
			for (int a = 0; a < NS; a++)
#pragma unroll
				for (int b = 0; b < NS; b++) {
					double s = 0.0;
#pragma unroll
					for (int r = 0; r <= NNS + (a < b ? a : b); r++) s += T[r][NNS + a] * T[r][NNS + b];
					As[a][b] = s;
				}
			if (sm) {  // A_s,b = A_s - delta z_s z_s^T / kappa
#pragma unroll
				for (int a = 0; a < NS; a++)
#pragma unroll
					for (int b = 0; b < NS; b++) As[a][b] -= delta * zu[NNS + a] * zu[NNS + b] / kappa;
			}
#pragma unroll
			for (int a = 0; a < NS; a++) ys[a] = au[NNS + a];
			if constexpr (NS == 1) {
				ys[0] = ys[0] / As[0][0];
			} else {
				const double det = As[0][0] * As[1][1] - As[0][1] * As[1][0];
				const double y0 = (As[1][1] * ys[0] - As[0][1] * ys[1]) / det;
				const double y1 = (As[0][0] * ys[1] - As[1][0] * ys[0]) / det;
				ys[0] = y0;
				ys[1] = y1;
			}
#pragma unroll
			for (int a = 0; a < NS; a++) ys[a] += bu[NNS + a];
			double es[N], ts[N];
#pragma unroll
			for (int r = 0; r < N; r++) {
				double s = 0.0;
				if (r < 6) {
#pragma unroll
					for (int a = 0; a < NS; a++)
						if (r <= NNS + a) s += T[r][NNS + a] * ys[a];
				}
				es[r] = s;
			}
			apply_q<N, 6, 0>(T, vhead, beta, es);
			mul_lower<N>(L, es, ts);
			// ---- joint strategy torques through the posture Jacobian: L^-1 J_post^T = Q [0; R'_22; 0] S_s^-1
			double u[N], vt[NS];
			double js[N];
			auto post_lambda = [&](double(&y)[NS]) {  // y <- Lambda_js,mod y,  A_js = D R'_22^T R'_22 D
				double zz[NS];
#pragma unroll
				for (int a = 0; a < NS; a++) {
					y[a] *= sig[NNS + a];  // D^-1
					zz[a] = 0.0;
				}
				// z_js = X_post^T g = D R'_22^T (Q^T g)[NNS:6]: expressed in the D^-1-scaled coordinates it is R'_22^T (Q^T g)[NNS:6]
				double qg[N];
				if (sm) {
#pragma unroll
					for (int r = 0; r < N; r++) qg[r] = g[r];
					apply_qt<N, 6, 0>(T, vhead, beta, qg);
#pragma unroll
					for (int a = 0; a < NS; a++) {
						double s = 0.0;
#pragma unroll
						for (int r = 0; r <= a; r++) s += T[NNS + r][NNS + a] * qg[NNS + r];
						zz[a] = s;
					}
				}
				solve_block_sm<N, NS, NNS>(T, rinv, y, zz, sm, delta, kappa);
#pragma unroll
				for (int a = 0; a < NS; a++) y[a] *= sig[NNS + a];
			};
			auto post_map = [&](const double(&y)[NS], double(&out)[N]) {  // out = J_post^T y = L Q [0; R'_22 D y; 0]
				double ee[N];
#pragma unroll
				for (int r = 0; r < N; r++) {
					double s = 0.0;
					if (r >= NNS && r < 6) {
#pragma unroll
						for (int a = 0; a < NS; a++)
							if (r - NNS <= a) s += T[r][NNS + a] * (y[a] / sig[NNS + a]);
					}
					ee[r] = s;
				}
				apply_q<N, 6, 0>(T, vhead, beta, ee);
				mul_lower<N>(L, ee, out);
			};
			if (c1 > c2 || p.enforce_type_1_strategy) {
				status |= OSC_STATUS_TYPE1;
#pragma unroll
				for (int j = 0; j < N; j++) u[j] = -p.kp_type_1 * (q[j] - q_prior[j]) - p.kv_type_1 * dq[j];
#pragma unroll
				for (int a = 0; a < NS; a++) {
					double s = 0.0;
#pragma unroll
					for (int j = 0; j < N; j++) s += Vs[j][a] * u[j];
					vt[a] = s;
				}
				post_lambda(vt);
				post_map(vt, js);
			} else {
				status |= OSC_STATUS_TYPE2;
				double dir[N];
#pragma unroll
				for (int j = 0; j < N; j++) {
					dir[j] = st[(int64_t)(MC_TYPE2_DIR + j) * NR + i];
					if (Vs[j][0] != 0.0) {
						if (fabs(q[j] - mdl.q_upper[j]) < p.type_2_angle_threshold)
							dir[j] = -1.0;
						else if (fabs(q[j] - mdl.q_lower[j]) < p.type_2_angle_threshold)
							dir[j] = 1.0;
					}
					st[(int64_t)(MC_TYPE2_DIR + j) * NR + i] = dir[j];
				}
				double nf = 0.0, fTd = 0.0;
#pragma unroll
				for (int k = 0; k < 6; k++) nf += (fstar[k] + F[k]) * (fstar[k] + F[k]);
				nf = sqrt(nf);
#pragma unroll
				for (int k = 0; k < 6; k++) fTd += (nf > 0.0 ? (fstar[k] + F[k]) / nf : (fstar[k] + F[k])) * Us[k][0];
#pragma unroll
				for (int j = 0; j < N; j++) u[j] = dir[j] * fabs(fTd) * p.type_2_torque_ratio * mdl.effort[j];
				double v1[NS], v2[NS], js2[N];
#pragma unroll
				for (int a = 0; a < NS; a++) {
					double s1 = 0.0, s2 = 0.0;
#pragma unroll
					for (int j = 0; j < N; j++) {
						s1 += Vs[j][a] * u[j];
						s2 += Vs[j][a] * (-p.kv_type_2 * dq[j]);
					}
					v1[a] = s1;
					v2[a] = s2;
				}
				post_map(v1, js);
				post_lambda(v2);
				post_map(v2, js2);
#pragma unroll
				for (int j = 0; j < N; j++) js[j] += js2[j];
			}
#pragma unroll
			for (int j = 0; j < N; j++) {
				double tsj = ts[j];
				if (isnan(tsj)) {
					tsj = 0.0;
					status |= OSC_STATUS_NAN_SCRUBBED;
				} else if (tsj > mdl.effort[j])
					tsj = mdl.effort[j];
				else if (tsj < -mdl.effort[j])
					tsj = -mdl.effort[j];
				tau[j] += alpha * tsj + (1.0 - alpha) * js[j];
			}
		}
	}

	// ---- JointTask in the null space N = N_js N_ns (or N_ns when NS == 0): complement of all six whitened columns
	if constexpr (HAS_JT) {
		constexpr int Mn = N - 6;
		if constexpr (Mn == 0) {
			status |= OSC_STATUS_ZERO_RANGE;
		} else {
			const DevJt& jt = P.jt[0];
			const osc_joint_params& jp = jt.p;
			double pid[N], acc[N];
			joint_control_law<N, N>(jt, NR, i, q, dq, pid, acc);
			double Qp[N][Mn], W[N][Mn], K[N][Mn];
#pragma unroll
			for (int a = 0; a < Mn; a++) {
				double ee[N];
#pragma unroll
				for (int j = 0; j < N; j++) ee[j] = (j == 6 + a) ? 1.0 : 0.0;
				apply_q<N, 6, 0>(T, vhead, beta, ee);
				double k[N];
				mul_lower<N>(L, ee, k);
#pragma unroll
				for (int j = 0; j < N; j++) Qp[j][a] = ee[j];
				solve_lower_t<N>(L, invd, ee);
#pragma unroll
				for (int j = 0; j < N; j++) {
					W[j][a] = ee[j];
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
			if (!(nrm_chk < 1.0e6)) status |= OSC_STATUS_UNHANDLED;
			cholesky_lower<Mn>(G, invg);
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
				double dj = 0.0, gj[N];
				int kj = 0;
#pragma unroll
				for (int j = 0; j < N; j++) {
					const double d = jp.bie_threshold - Mdiag[j];
					gj[j] = (d > 0.0) ? 1.0 : 0.0;
					dj += (d > 0.0) ? d : 0.0;
					kj += (d > 0.0) ? 1 : 0;
				}
				if (kj >= 2) status |= OSC_STATUS_UNHANDLED;  // state is already updated: cannot fall back any more (checked up front below)
				if (kj == 1) {
					solve_lower<N>(L, invd, gj);
					double muj = 0.0;
#pragma unroll
					for (int j = 0; j < N; j++) muj += gj[j] * gj[j];
					double c[Mn], cc = 0.0, cz = 0.0;
#pragma unroll
					for (int a = 0; a < Mn; a++) {
						double s = 0.0;
#pragma unroll
						for (int j = 0; j < N; j++) s += Qp[j][a] * gj[j];
						c[a] = s;
						cc += s * s;
						cz += s * z2[a];
					}
					const double f = dj * cz / ((1.0 + dj * muj) - dj * cc);
#pragma unroll
					for (int a = 0; a < Mn; a++) z2[a] += c[a] * f;
				}
			}
#pragma unroll
			for (int j = 0; j < N; j++) {
				double s = 0.0;
#pragma unroll
				for (int a = 0; a < Mn; a++) s += K[j][a] * (z1[a] + z2[a]);
				tau[j] += s;
			}
		}
	}

	if (P.torque_saturation) {
#pragma unroll
		for (int j = 0; j < N; j++) tau[j] = fmin(fmax(tau[j], -mdl.effort[j]), mdl.effort[j]);
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
	return true;
}

// One robot of the hand-over list: returns false when the robot needs the general path.
template <int N, bool HAS_JT>
DEVI bool blend_robot(const OscProgram& P, const int64_t i) {
	const int64_t NR = P.n_robots;
	const DevModel& mdl = P.model;
	const DevMft& t = P.mft[0];
	const osc_mft_params& p = t.p;
	// cases this file does not specialise
	if (!t.full || t.rank != 6) return false;
	if (HAS_JT && p.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES && P.jt[0].p.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES &&
		P.jt[0].p.bie_threshold != p.bie_threshold)
		return false;

	double q[N], dq[N];
#pragma unroll
	for (int j = 0; j < N; j++) {
		q[j] = P.q[(int64_t)j * NR + i];
		dq[j] = P.dq[(int64_t)j * NR + i];
	}
	KinDyn<N> kd;
	forward_kinematics<N>(mdl, q, kd);
	mass_matrix<N, true>(mdl, kd);
	double x[3], Rc[9];
	frame_pose<N>(kd, t.body, t.ctrl_R, t.ctrl_t, x, Rc);
	double JT0[N][6];
	point_jacobian_t<N>(mdl, kd, t.body, x, JT0);
	double G[6][6], U[6][6];
#pragma unroll
	for (int a = 0; a < 6; a++)
#pragma unroll
		for (int b = 0; b <= a; b++) {
			double s = 0.0;
#pragma unroll
			for (int j = 0; j < N; j++) s += JT0[j][a] * JT0[j][b];
			G[a][b] = s;
			G[b][a] = s;
		}
	jacobi_eig6(G, U);
	// sort eigenpairs by decreasing eigenvalue (selection by compare-exchange: static register indices only)
	double lam[6];
#pragma unroll
	for (int a = 0; a < 6; a++) lam[a] = G[a][a];
#pragma unroll
	for (int a = 0; a < 5; a++)
#pragma unroll
		for (int b = a + 1; b < 6; b++) {
			if (lam[b] > lam[a]) {
				const double tl = lam[a];
				lam[a] = lam[b];
				lam[b] = tl;
#pragma unroll
				for (int k = 0; k < 6; k++) {
					const double tu = U[k][a];
					U[k][a] = U[k][b];
					U[k][b] = tu;
				}
			}
		}
	double sig[6];
#pragma unroll
	for (int a = 0; a < 6; a++) sig[a] = sqrt(fmax(lam[a], 0.0));
	if (sig[0] < p.s_abs_tol) return false;	 // fully singular task
	int n_ns = 6;
	double alpha = 1.0;
#pragma unroll
	for (int c = 5; c >= 1; c--) {
		const double icn = sig[c] / sig[0];
		if (icn < p.s_max) {
			n_ns = c;
			alpha = fmin(fmax((icn - p.s_min) / (p.s_max - p.s_min), 0.0), 1.0);
		}
	}
	const int n_s = 6 - n_ns;
	if (n_s > 2) return false;
	// also to the general path: several clamped inertia entries / handling off with singular directions (decided before any state is touched)
	{
		int k1 = 0, k2 = 0;
#pragma unroll
		for (int j = 0; j < N; j++) {
			k1 += (p.bie_threshold - kd.M[j][j] > 0.0) ? 1 : 0;
			if (HAS_JT) k2 += (P.jt[0].p.bie_threshold - kd.M[j][j] > 0.0) ? 1 : 0;
		}
		if (p.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES && k1 >= 2) return false;
		if (HAS_JT && P.jt[0].p.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES && k2 >= 2) return false;
		if (n_s > 0 && !p.singularity_handling_enabled) return false;
	}

	// dynamics factor and whitened Jacobian
	double Mdiag[N], invd[N], L[N][N];
#pragma unroll
	for (int r = 0; r < N; r++) {
		Mdiag[r] = kd.M[r][r];
#pragma unroll
		for (int c = 0; c <= r; c++) L[r][c] = kd.M[r][c];
	}
	cholesky_lower<N>(L, invd);
	double Xw[N][6];
#pragma unroll
	for (int r = 0; r < N; r++)
#pragma unroll
		for (int a = 0; a < 6; a++) {
			double s = JT0[r][a];
#pragma unroll
			for (int k = 0; k < r; k++) s -= L[r][k] * Xw[k][a];
			Xw[r][a] = s * invd[r];
		}
	if (n_s == 0) return blend_path<N, 0, HAS_JT>(P, i, q, dq, kd, L, invd, Mdiag, JT0, Xw, U, sig, alpha, x, Rc, OSC_STATUS_SINGULAR_PATH);
	if (n_s == 1) return blend_path<N, 1, HAS_JT>(P, i, q, dq, kd, L, invd, Mdiag, JT0, Xw, U, sig, alpha, x, Rc, OSC_STATUS_SINGULAR_PATH);
	return blend_path<N, 2, HAS_JT>(P, i, q, dq, kd, L, invd, Mdiag, JT0, Xw, U, sig, alpha, x, Rc, OSC_STATUS_SINGULAR_PATH);
}

// Hand-over list of the fast kernel for the flagship hierarchy: unrolled blending path, general path as fallback.
template <int N, bool HAS_JT>
__global__ void __launch_bounds__(64) osc_blend_kernel(const __grid_constant__ OscProgram P) {
	// let the next cycle's fast kernel get scheduled behind this one right away, then wait for the fast kernel of this cycle
	asm volatile("griddepcontrol.launch_dependents;");
	asm volatile("griddepcontrol.wait;" ::: "memory");
	const int32_t count = P.sing_count[P.sing_parity];
	const int stride = gridDim.x * blockDim.x;
	for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < count; slot += stride) {
		const int64_t i = (int64_t)P.sing_list[slot];
		if (!blend_robot<N, HAS_JT>(P, i)) generic_cycle_one<N>(P, i, OSC_STATUS_SINGULAR_PATH);
	}
	publish_general_done(P, count);
}

}  // namespace osc
