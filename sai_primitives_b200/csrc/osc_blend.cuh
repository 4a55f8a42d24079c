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
// eigen-decomposition J J^T = U S^2 U^T (6 x 6 Gram matrix: tridiagonalisation + implicit QL, osc_eig6.h):
//   U = [U_ns | U_s] (first NNS / last NS columns),  V_s = J^T U_s S_s^-1 = L X U_s S_s^-1
//   X U = Q R' (Householder):  J_ns M^-1 J_ns^T = R'_11^T R'_11,   J_s M^-1 J_s^T = R'_12^T R'_12 + R'_22^T R'_22,
//   posture Jacobian J_post = V_s^T N_ns:  L^-1 J_post^T = Q [0; R'_22; 0] S_s^-1,  so
//   J_post M^-1 J_post^T = S_s^-1 R'_22^T R'_22 S_s^-1,  and N = N_js N_ns is L^-T (I - Q_6 Q_6^T) L^T:
//   the joint task sees exactly the complement Q e_7 of the fast path.
#pragma once
#include "osc_kindyn.cuh"
#include "osc_singular.cuh"
#include "osc_tasks.cuh"
#define OSC_EIG6_LEAN_MATH
#include "osc_eig6.h"
#include <type_traits>

namespace osc {

// (A - delta z z^T / kappa)^-1 y for A = Rb^T Rb given by its upper-triangular factor stored in T[C0 + i][C0 + j] (i <= j),
// size S; rinv = reciprocal diagonal.  use_sm = false gives plain A^-1 y.
template <int N, int S, int C0, class TT>
DEVI void solve_block_sm(const TT& T, const double (&rinv)[6], double (&y)[S], const double (&z)[S], bool use_sm, double delta,
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

// Where blend_path reads the results of the first half from: the registers of the same thread (osc_blend_kernel) or the
// scratch block of the split path (BlendLayout below), element by element at the point of use -- the 150 doubles of L, J, X
// and U are then never live at the same time as the control law and the classification, which is what spilled.
template <int N>
struct BlendRegisterSource {
	const double (&grav_)[N];
	const double (&L_)[N][N];
	const double (&invd_)[N];
	const double (&Mdiag_)[N];
	const double (&JT0_)[N][6];
	const double (&U_)[6][6];
	DEVI double grav(int j) const { return grav_[j]; }
	DEVI double l(int r, int c) const { return L_[r][c]; }
	DEVI double invd(int r) const { return invd_[r]; }
	DEVI double mdiag(int r) const { return Mdiag_[r]; }
	DEVI double jt0(int j, int a) const { return JT0_[j][a]; }
	DEVI double u(int a, int c) const { return U_[a][c]; }
};
template <int N>
struct BlendScratchSource {
	using BL = BlendLayout<N>;
	const double* S;  // scratch + slot
	int64_t cap;
	DEVI double at(int c) const { return S[(int64_t)c * cap]; }
	DEVI double grav(int j) const { return at(BL::GRAV + j); }
	DEVI double l(int r, int c) const { return at(BL::LL + r * (r + 1) / 2 + c); }
	DEVI double invd(int r) const { return at(BL::INVD + r); }
	DEVI double mdiag(int r) const { return at(BL::MDIAG + r); }
	DEVI double jt0(int j, int a) const { return at(BL::JT0 + j * 6 + a); }
	DEVI double u(int a, int c) const { return at(BL::U + a * 6 + c); }
};

// Everything after the eigen-decomposition, for NS singular directions (NS = 0: the thin band of non-singular robots
// the sound test of the fast kernel rejects).  Returns false when the case must go to the general path.
// MOTION: the task is under pure motion control (host check, as for the fused kernel): only the two PID laws are compiled in and F = 0.
// SMEM: the factor L and the matrix T (70 of the doubles with the longest lives) are kept in shared memory, element e of
// this thread at smem[e * kBlendBlock], instead of registers (where, with everything else, they spill).
constexpr int kBlendBlock = 64;
// measured: with L and T in shared memory the variants kernel is 2 % slower (143 KB of shared memory per SM leave the spills
// 100 KB of L1 instead of 240 KB), so the split path runs with SMEM off; the switch stays for the next restructuring
constexpr bool kBlendSmem = false;
template <int N>
constexpr int blend_smem_doubles() {
	return N * (N + 1) / 2 + 6 * N;
}
template <int N, int NS, bool HAS_JT, class Src, bool MOTION = false, bool SMEM = false>
DEVI bool blend_path(const OscProgram& P, int64_t i, const double (&q)[N], const double (&dq)[N], const Src& src, const double (&sig)[6], double alpha,
					 const double x[3], const double Rc[9], uint32_t status, double* smem = nullptr) {
	constexpr int NNS = 6 - NS;
	const int64_t NR = P.n_robots;
	const DevModel& mdl = P.model;
	const DevMft& t = P.mft[0];
	const osc_mft_params& p = t.p;
	double* st = t.st;
	int32_t* ist = t.ist;
	const int dec = p.dynamic_decoupling_type;

	// bounded inertia: one clamped entry -> rank-one (Sherman-Morrison); several -> general path
	double delta = 0.0;
	int kclamp = 0;
	if (dec == OSC_BOUNDED_INERTIA_ESTIMATES) {
#pragma unroll
		for (int j = 0; j < N; j++) {
			const double dj = p.bie_threshold - src.mdiag(j);
			delta += (dj > 0.0) ? dj : 0.0;
			kclamp += (dj > 0.0) ? 1 : 0;
		}
		if (kclamp >= 2) return false;
	}
	const bool sm = (dec == OSC_BOUNDED_INERTIA_ESTIMATES) && kclamp == 1;

	// right singular vectors of the singular directions: V_s = J^T U_s / sigma, oriented so that the largest-magnitude
	// entry is positive (sign convention of this repo); the matching U_s column follows the sign
	double Us[6][NS > 0 ? NS : 1], Vs[N][NS > 0 ? NS : 1];
	double sgn_s[NS > 0 ? NS : 1];	// [U_ns | U_s] with the final signs is  src.u(a, c) * (c < NNS ? 1 : sgn_s[c - NNS])
	if constexpr (NS > 0) {
#pragma unroll
		for (int c = 0; c < NS; c++) {
			double vmax = 0.0, amax = -1.0;
			const double inv = 1.0 / sig[NNS + c];
#pragma unroll
			for (int a = 0; a < 6; a++) Us[a][c] = src.u(a, NNS + c);
#pragma unroll
			for (int j = 0; j < N; j++) {
				double s = 0.0;
#pragma unroll
				for (int a = 0; a < 6; a++) s += src.jt0(j, a) * Us[a][c];
				s *= inv;
				Vs[j][c] = s;
				if (fabs(s) > amax) {
					amax = fabs(s);
					vmax = s;
				}
			}
			const double sgn = (vmax < 0.0) ? -1.0 : 1.0;
			sgn_s[c] = sgn;
#pragma unroll
			for (int j = 0; j < N; j++) Vs[j][c] *= sgn;
#pragma unroll
			for (int a = 0; a < 6; a++) Us[a][c] = sgn * Us[a][c];
		}
	}

	// task velocity J0 dq
	double v[3] = {0, 0, 0}, w[3] = {0, 0, 0};
#pragma unroll
	for (int j = 0; j < N; j++)
#pragma unroll
		for (int k = 0; k < 3; k++) {
			v[k] += src.jt0(j, k) * dq[j];
			w[k] += src.jt0(j, 3 + k) * dq[j];
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
	mft_control_law<MOTION>(t, NR, i, x, Rc, v, w, P.write_observers != 0, fstar, F, status);

	// dynamics factor (read only now: nothing above needs it), and the unit vector of the clamped entry through it
	typename std::conditional<SMEM, SmLowerMat<kBlendBlock>, RegMat<N, N>>::type L;
	typename std::conditional<SMEM, SmMat<6, kBlendBlock>, RegMat<N, 6>>::type T;
	if constexpr (SMEM) {
		L.b = smem;
		T.b = smem + (N * (N + 1) / 2) * kBlendBlock;
	}
	double invd[N], g[N], mu = 0.0;
#pragma unroll
	for (int r = 0; r < N; r++) {
		invd[r] = src.invd(r);
#pragma unroll
		for (int c = 0; c <= r; c++) L[r][c] = src.l(r, c);
	}
	if (sm) {
#pragma unroll
		for (int j = 0; j < N; j++) g[j] = (p.bie_threshold - src.mdiag(j) > 0.0) ? 1.0 : 0.0;
		solve_lower<N>(L, invd, g);
#pragma unroll
		for (int j = 0; j < N; j++) mu += g[j] * g[j];
	}
	const double kappa = 1.0 + delta * mu;
	// ---- T = X U with X = L^-1 J^T (its Householder QR follows), and the coordinates of f*, F in the rotated task basis
	double au[6], bu[6];
	{
		double Ue[6][6];
#pragma unroll
		for (int a = 0; a < 6; a++)
#pragma unroll
			for (int c = 0; c < 6; c++) Ue[a][c] = (NS > 0 && c >= NNS) ? sgn_s[c >= NNS ? c - NNS : 0] * src.u(a, c) : src.u(a, c);
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
		// T = X U = L^-1 (J^T U): row r of J^T U, then the forward substitution with the rows above it
#pragma unroll
		for (int r = 0; r < N; r++) {
			double jr[6];
#pragma unroll
			for (int a = 0; a < 6; a++) jr[a] = src.jt0(r, a);
#pragma unroll
			for (int c = 0; c < 6; c++) {
				double s = 0.0;
#pragma unroll
				for (int a = 0; a < 6; a++) s += jr[a] * Ue[a][c];
#pragma unroll
				for (int k = 0; k < r; k++) s -= L[r][k] * T[k][c];
				T[r][c] = s * invd[r];
			}
		}
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
	double vhead[6], beta[6], rinv[6];
	householder_qr<N, 6, 0>(T, vhead, beta, rinv);

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
					const double d = jp.bie_threshold - src.mdiag(j);
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
		for (int j = 0; j < N; j++) tau[j] += src.grav(j);
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
	// sort eigenpairs by decreasing eigenvalue (selection by compare-exchange: static register indices only)
	double lam[6];
	sym_eig6(G, U, lam);
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

	// dynamics factor
	double Mdiag[N], invd[N], L[N][N];
#pragma unroll
	for (int r = 0; r < N; r++) {
		Mdiag[r] = kd.M[r][r];
#pragma unroll
		for (int c = 0; c <= r; c++) L[r][c] = kd.M[r][c];
	}
	cholesky_lower<N>(L, invd);
	const BlendRegisterSource<N> src{kd.g, L, invd, Mdiag, JT0, U};
	if (n_s == 0) return blend_path<N, 0, HAS_JT>(P, i, q, dq, src, sig, alpha, x, Rc, OSC_STATUS_SINGULAR_PATH);
	if (n_s == 1) return blend_path<N, 1, HAS_JT>(P, i, q, dq, src, sig, alpha, x, Rc, OSC_STATUS_SINGULAR_PATH);
	return blend_path<N, 2, HAS_JT>(P, i, q, dq, src, sig, alpha, x, Rc, OSC_STATUS_SINGULAR_PATH);
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

// ---------------------------------------------------------------------------------------------------------------------------
// The same path split over kernels, for cycles in which the host hint (osc_pipeline.cuh) says that many robots are on the
// general path.  osc_blend_kernel above recomputes kinematics and dynamics, and walks the eigen-decomposition and up to
// three inlined variants (0 / 1 / 2 singular directions) in one 1 MB instruction stream; a warp with mixed robots walks
// the variants one after the other.  Here
//   0. the fused kernel parks what it had already computed when it decided to hand the robot over (pose, M = L L^T, the
//      Jacobian: park_for_blend, osc_cycle.cuh) in a scratch block, component c of list slot s at
//      scratch[c * cap + s];
//   1. osc_blend_classify_kernel does the eigen-decomposition and the classification and appends the slot to the list of
//      its variant (warp-aggregated atomics: runs of consecutive slots);
//   2. osc_blend_variants_kernel deals the three lists to its blocks chunk by chunk, so that a warp only ever executes one
//      variant, reading the parked state element by element where blend_path uses it;
//   3. osc_blend_fallback_kernel takes what the blending path does not specialise through the rolled general path and
//      publishes the end of the cycle's general path.
// Classification of list slot `slot` from what the fused kernel parked there: returns the variant (number of singular
// directions, 0..2) or 3 when the robot needs the general path (the tests of blend_robot, in the same order).
template <int N, bool HAS_JT>
DEVI int blend_classify(const OscProgram& P, const int64_t slot) {
	using BL = BlendLayout<N>;
	const int64_t cap = P.blend_cap;
	double* S = P.blend_scratch + slot;
	const DevMft& t = P.mft[0];
	const osc_mft_params& p = t.p;
	if (!t.full || t.rank != 6) return 3;
	if (HAS_JT && p.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES && P.jt[0].p.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES &&
		P.jt[0].p.bie_threshold != p.bie_threshold)
		return 3;
	double G[6][6], U[6][6], lam[6];
	{
		double JT0[N][6];
#pragma unroll
		for (int j = 0; j < N; j++)
#pragma unroll
			for (int a = 0; a < 6; a++) JT0[j][a] = S[(int64_t)(BL::JT0 + j * 6 + a) * cap];
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
	}
	int k1 = 0, k2 = 0;
#pragma unroll
	for (int r = 0; r < N; r++) {
		const double m = S[(int64_t)(BL::MDIAG + r) * cap];
		k1 += (p.bie_threshold - m > 0.0) ? 1 : 0;
		if (HAS_JT) k2 += (P.jt[0].p.bie_threshold - m > 0.0) ? 1 : 0;
	}
	sym_eig6(G, U, lam);
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
	if (sig[0] < p.s_abs_tol) return 3;	 // fully singular task
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
	if (n_s > 2) return 3;
	if (p.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES && k1 >= 2) return 3;
	if (HAS_JT && P.jt[0].p.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES && k2 >= 2) return 3;
	if (n_s > 0 && !p.singularity_handling_enabled) return 3;
#pragma unroll
	for (int a = 0; a < 6; a++) {
		S[(int64_t)(BL::SIG + a) * cap] = sig[a];
#pragma unroll
		for (int b = 0; b < 6; b++) S[(int64_t)(BL::U + a * 6 + b) * cap] = U[a][b];
	}
	S[(int64_t)BL::ALPHA * cap] = alpha;
	return n_s;
}

// The state parked by the fused kernel and the classification kernel, then blend_path of the variant.
template <int N, int NS, bool HAS_JT, bool MOTION>
DEVI void blend_variant(const OscProgram& P, const int64_t i, const int64_t slot, double* sm) {
	using BL = BlendLayout<N>;
	const int64_t cap = P.blend_cap;
	const BlendScratchSource<N> src{P.blend_scratch + slot, cap};
	double q[N], dq[N], sig[6], x[3], Rc[9];
#pragma unroll
	for (int r = 0; r < N; r++) {
		q[r] = src.at(BL::Q + r);
		dq[r] = src.at(BL::DQ + r);
	}
#pragma unroll
	for (int a = 0; a < 6; a++) sig[a] = src.at(BL::SIG + a);
#pragma unroll
	for (int k = 0; k < 3; k++) x[k] = src.at(BL::X + k);
#pragma unroll
	for (int k = 0; k < 9; k++) Rc[k] = src.at(BL::RC + k);
	const double alpha = src.at(BL::ALPHA);
	if (!blend_path<N, NS, HAS_JT, BlendScratchSource<N>, MOTION, kBlendSmem>(P, i, q, dq, src, sig, alpha, x, Rc, OSC_STATUS_SINGULAR_PATH, sm)) {
		// cannot happen: blend_classify sends every case blend_path refuses to the general path before any state is touched
		const int64_t NR = P.n_robots;
#pragma unroll
		for (int j = 0; j < N; j++) P.tau[(int64_t)j * NR + i] = __longlong_as_double(0x7ff8000000000000LL);
		P.status[i] = OSC_STATUS_SINGULAR_PATH | OSC_STATUS_UNHANDLED;
	}
}

template <int N, bool HAS_JT>
__global__ void __launch_bounds__(64) osc_blend_classify_kernel(const __grid_constant__ OscProgram P) {
	asm volatile("griddepcontrol.launch_dependents;");
	asm volatile("griddepcontrol.wait;" ::: "memory");
	// the variant lists were cleared by the last general-path kernel of the previous cycle
	if (P.general_done) {
		if ((threadIdx.x & 31) == 0)
			while ((int32_t)(ld_acquire_u32(P.general_done) - (P.epoch - 1u)) < 0) __nanosleep(128);
		__syncwarp();
	}
	const int32_t count = P.sing_count[P.sing_parity];
	const int stride = gridDim.x * blockDim.x;
	const int lane = threadIdx.x & 31;
	for (int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < count; base += stride) {
		const int slot = base + lane;
		int variant = -1;
		int32_t robot = 0;
		if (slot < count) {
			robot = P.sing_list[slot];
			variant = blend_classify<N, HAS_JT>(P, (int64_t)slot);
		}
#pragma unroll
		for (int v = 0; v < 4; v++) {
			const unsigned m = __ballot_sync(0xffffffffu, variant == v);
			if (m) {
				const int leader = __ffs(m) - 1;
				int32_t at = 0;
				if (lane == leader) at = atomicAdd(&P.blend_counts[v], __popc(m));
				at = __shfl_sync(0xffffffffu, at, leader);
				if (variant == v) P.blend_lists[(int64_t)v * P.blend_cap + at + __popc(m & ((1u << lane) - 1u))] = (v == 3) ? robot : slot;
			}
		}
	}
}

// The three variant lists are cut into chunks of one block each and the chunks of all three dealt to the blocks together:
// a block (hence a warp) only ever executes one variant at a time, and the three lists are worked on at the same time.
template <int N, int NS, bool HAS_JT, bool MOTION>
DEVI void blend_variant_chunk(const OscProgram& P, int chunk, int32_t count, double* sm) {
	const int k = chunk * (int)blockDim.x + (int)threadIdx.x;
	if (k < count) {
		const int32_t slot = P.blend_lists[(int64_t)NS * P.blend_cap + k];
		blend_variant<N, NS, HAS_JT, MOTION>(P, (int64_t)P.sing_list[slot], (int64_t)slot, sm);
	}
}
// 255 registers, four blocks per SM: capping the registers for six or eight blocks costs 14 % / 21 % (profiles/r02_summary.md)
template <int N, bool HAS_JT, bool MOTION>
__global__ void __launch_bounds__(kBlendBlock) osc_blend_variants_kernel(const __grid_constant__ OscProgram P) {
	asm volatile("griddepcontrol.launch_dependents;");
	asm volatile("griddepcontrol.wait;" ::: "memory");
	extern __shared__ double blend_sm[];  // blend_smem_doubles<N>() per thread when kBlendSmem
	double* sm = kBlendSmem ? blend_sm + threadIdx.x : nullptr;
	const int32_t c0 = P.blend_counts[0], c1 = P.blend_counts[1], c2 = P.blend_counts[2];
	const int bs = (int)blockDim.x;
	const int b1 = (c1 + bs - 1) / bs, b2 = (c2 + bs - 1) / bs, b0 = (c0 + bs - 1) / bs;
	// Chunk w of the concatenated lists goes to block w mod gridDim.x, the most common variant (one singular direction) first: at
	// any time nearly all blocks of an SM execute the same variant, which is what keeps its 140 KB instruction stream
	// cached.  Longest-first and work-queue orders were measured: both lose 5-15 % beyond one wave (profiles/r02_summary.md).
	for (int w = blockIdx.x; w < b1 + b2 + b0; w += gridDim.x) {
		if (w < b1)
			blend_variant_chunk<N, 1, HAS_JT, MOTION>(P, w, c1, sm);
		else if (w < b1 + b2)
			blend_variant_chunk<N, 2, HAS_JT, MOTION>(P, w - b1, c2, sm);
		else
			blend_variant_chunk<N, 0, HAS_JT, MOTION>(P, w - b1 - b2, c0, sm);
	}
}

template <int N>
__global__ void __launch_bounds__(64, OSC_GENERIC_MIN_BLOCKS) osc_blend_fallback_kernel(const __grid_constant__ OscProgram P) {
	asm volatile("griddepcontrol.launch_dependents;");
	asm volatile("griddepcontrol.wait;" ::: "memory");
	const int32_t count = P.blend_counts[3];
	const int32_t* list = P.blend_lists + (int64_t)3 * P.blend_cap;
	const int stride = gridDim.x * blockDim.x;
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) generic_cycle_one<N>(P, (int64_t)list[k], OSC_STATUS_SINGULAR_PATH);
	publish_general_done(P, P.sing_count[P.sing_parity], P.blend_counts);
}

}  // namespace osc
