// The singular / blending path: robots whose task Jacobian fails the sound non-singularity test of the
// fast kernel are appended to a compacted list and re-evaluated here, one robot per thread, following the
// reference statement by statement with an explicit thin SVD (from the eigen-decomposition of the Gram matrix, osc_eig6.h):
//   SingularityHandler::updateTaskModel     src/tasks/SingularityHandler.cpp:75-228
//   SingularityHandler::classifySingularity src/tasks/SingularityHandler.cpp:230-295 (stateful)
//   SingularityHandler::computeTorques      src/tasks/SingularityHandler.cpp:297-368
//   JointTask::updateTaskModel/computeTorques with a general N_prec  src/tasks/JointTask.cpp:218-356
// Loops are rolled and matrices live in thread-local memory: this path trades speed for generality (it also
// serves the thin band of non-singular robots the sound test rejects, so it implements every branch).
#pragma once
#include "osc_kindyn.cuh"
#include "osc_pipeline.cuh"
#include "osc_tasks.cuh"
#define OSC_EIG6_LEAN_MATH
#include "osc_eig6.h"

#ifndef OSC_GENERIC_MIN_BLOCKS
#define OSC_GENERIC_MIN_BLOCKS 4
#endif

namespace osc {
namespace sg {

constexpr int MAXD = OSC_MAX_DOF;  // 8

// ---- tiny dense helpers on row-major buffers with runtime sizes (rolled loops on purpose)
// (two output columns per pass: two independent accumulation chains, half the loop overhead -- the general path runs with
// one or two warps per scheduler and is bound by the latency of exactly these chains)
static __device__ __noinline__ void mm(const double* A, int ar, int ac, const double* B, int bc, double* C) {	// C = A B
	for (int i = 0; i < ar; i++) {
		int j = 0;
		for (; j + 1 < bc; j += 2) {
			double s0 = 0.0, s1 = 0.0;
			for (int k = 0; k < ac; k++) {
				const double a = A[i * ac + k];
				s0 += a * B[k * bc + j];
				s1 += a * B[k * bc + j + 1];
			}
			C[i * bc + j] = s0;
			C[i * bc + j + 1] = s1;
		}
		if (j < bc) {
			double s = 0.0;
			for (int k = 0; k < ac; k++) s += A[i * ac + k] * B[k * bc + j];
			C[i * bc + j] = s;
		}
	}
}
static __device__ __noinline__ void mm_at(const double* A, int ar, int ac, const double* B, int bc, double* C) {  // C = A^T B  (A: ar x ac, B: ar x bc)
	for (int i = 0; i < ac; i++) {
		int j = 0;
		for (; j + 1 < bc; j += 2) {
			double s0 = 0.0, s1 = 0.0;
			for (int k = 0; k < ar; k++) {
				const double a = A[k * ac + i];
				s0 += a * B[k * bc + j];
				s1 += a * B[k * bc + j + 1];
			}
			C[i * bc + j] = s0;
			C[i * bc + j + 1] = s1;
		}
		if (j < bc) {
			double s = 0.0;
			for (int k = 0; k < ar; k++) s += A[k * ac + i] * B[k * bc + j];
			C[i * bc + j] = s;
		}
	}
}
static __device__ __noinline__ void mm_bt(const double* A, int ar, int ac, const double* B, int br, double* C) {  // C = A B^T  (B: br x ac)
	for (int i = 0; i < ar; i++) {
		int j = 0;
		for (; j + 1 < br; j += 2) {
			double s0 = 0.0, s1 = 0.0;
			for (int k = 0; k < ac; k++) {
				const double a = A[i * ac + k];
				s0 += a * B[j * ac + k];
				s1 += a * B[(j + 1) * ac + k];
			}
			C[i * br + j] = s0;
			C[i * br + j + 1] = s1;
		}
		if (j < br) {
			double s = 0.0;
			for (int k = 0; k < ac; k++) s += A[i * ac + k] * B[j * ac + k];
			C[i * br + j] = s;
		}
	}
}
static __device__ __noinline__ void mv(const double* A, int ar, int ac, const double* x, double* y) {
	for (int i = 0; i < ar; i++) {
		double s = 0.0;
		for (int k = 0; k < ac; k++) s += A[i * ac + k] * x[k];
		y[i] = s;
	}
}
static __device__ __noinline__ void mv_t(const double* A, int ar, int ac, const double* x, double* y) {  // y = A^T x
	for (int j = 0; j < ac; j++) {
		double s = 0.0;
		for (int k = 0; k < ar; k++) s += A[k * ac + j] * x[k];
		y[j] = s;
	}
}
static __device__ __noinline__ double fro2(const double* A, int n) {
	double s = 0.0;
	for (int i = 0; i < n; i++) s += A[i] * A[i];
	return s;
}
// inverse of a symmetric positive definite matrix through its Cholesky factor (the reference uses
// .inverse() / .llt().solve(I) on such matrices); returns false when not positive definite
static __device__ __noinline__ bool spd_inverse(const double* A, int n, double* Ainv) {
	double L[MAXD * MAXD], invl[MAXD];	// one division per pivot; the substitutions multiply by the reciprocal
	bool ok = true;
	for (int j = 0; j < n; j++) {
		double d = A[j * n + j];
		for (int k = 0; k < j; k++) d -= L[j * n + k] * L[j * n + k];
		ok = ok && (d > 0.0);
		const double l = sqrt(d);
		const double il = 1.0 / l;
		L[j * n + j] = l;
		invl[j] = il;
		for (int i = j + 1; i < n; i++) {
			double s = A[i * n + j];
			for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
			L[i * n + j] = s * il;
		}
	}
	for (int c = 0; c < n; c++) {
		double x[MAXD];
		for (int i = 0; i < c; i++) x[i] = 0.0;	 // column c of L^-1 starts at row c
		for (int i = c; i < n; i++) {
			double s = (i == c) ? 1.0 : 0.0;
			for (int k = c; k < i; k++) s -= L[i * n + k] * x[k];
			x[i] = s * invl[i];
		}
		for (int i = n - 1; i >= 0; i--) {
			double s = x[i];
			for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
			x[i] = s * invl[i];
		}
		for (int i = 0; i < n; i++) Ainv[i * n + c] = x[i];
	}
	return ok;
}

// Thin SVD of the six-row projected task Jacobian A (6 x n, row-major), k = min(6, n):  A = U diag(s) V^T through the
// eigen-decomposition of the 6 x 6 Gram matrix A A^T in registers (osc_eig6.h), V = A^T U / s -- what the blending path does
// (osc_blend.cuh), a tenth of the instructions of the one-sided Jacobi sweeps in local memory that did this until round 2 (a third of the
// rolled general path: profiles/r02_summary.md).  Same conventions: s descending, a zero singular value gets a zero column of V, the
// largest-magnitude entry of every column of V is positive.
static __device__ __noinline__ void svd_rows6(const double* A, int n, double* U, double* s, double* V) {
	const int k = n < 6 ? n : 6;
	double G[6][6], Z[6][6], d[6];
#pragma unroll
	for (int a = 0; a < 6; a++)
#pragma unroll
		for (int b = 0; b <= a; b++) {
			double t = 0.0;
			for (int j = 0; j < n; j++) t += A[a * n + j] * A[b * n + j];
			G[a][b] = t;
			G[b][a] = t;
		}
	sym_eig6(G, Z, d);
#pragma unroll
	for (int a = 0; a < 5; a++)
#pragma unroll
		for (int b = a + 1; b < 6; b++) {
			if (d[b] > d[a]) {
				const double td = d[a];
				d[a] = d[b];
				d[b] = td;
#pragma unroll
				for (int r = 0; r < 6; r++) {
					const double tz = Z[r][a];
					Z[r][a] = Z[r][b];
					Z[r][b] = tz;
				}
			}
		}
#pragma unroll
	for (int jj = 0; jj < 6; jj++) {
		if (jj < k) {
			const double sv = sqrt(fmax(d[jj], 0.0));
			s[jj] = sv;
			const double inv = sv > 0.0 ? 1.0 / sv : 0.0;
#pragma unroll
			for (int a = 0; a < 6; a++) U[a * k + jj] = Z[a][jj];
			int im = 0;
			double vm = 0.0;
			for (int i = 0; i < n; i++) {
				double t = 0.0;
#pragma unroll
				for (int a = 0; a < 6; a++) t += A[a * n + i] * Z[a][jj];
				t *= inv;
				V[i * k + jj] = t;
				if (fabs(t) > fabs(vm)) {
					vm = t;
					im = i;
				}
			}
			(void)im;
			if (vm < 0.0) {
				for (int i = 0; i < n; i++) V[i * k + jj] = -V[i * k + jj];
#pragma unroll
				for (int a = 0; a < 6; a++) U[a * k + jj] = -U[a * k + jj];
			}
		}
	}
}

// Left singular vectors (U, k x k, row-major) and singular values (s, descending) of the joint task's projected Jacobian A
// (k x n, k <= N) for SaiModel::matrixRangeBasis (JointTask.cpp:218-283): eigen-decomposition of the k x k Gram matrix A A^T in
// registers, padded to N x N with -1 on the diagonal so that the padding sorts last.  The range basis only enters through
// U^T (.) U products, so any orthonormal basis of the same subspace gives the same torques.
template <int N>
static __device__ __noinline__ void left_svd_gram(const double* A, int k, int n, double* U, double* s) {
	double G[N][N], Z[N][N], d[N];
#pragma unroll
	for (int a = 0; a < N; a++)
#pragma unroll
		for (int b = 0; b <= a; b++) {
			double t = (a == b) ? -1.0 : 0.0;
			if (a < k && b < k) {
				t = 0.0;
				for (int j = 0; j < n; j++) t += A[a * n + j] * A[b * n + j];
			}
			G[a][b] = t;
			G[b][a] = t;
		}
	sym_eig<N>(G, Z, d);
#pragma unroll
	for (int a = 0; a < N - 1; a++)
#pragma unroll
		for (int b = a + 1; b < N; b++) {
			if (d[b] > d[a]) {
				const double td = d[a];
				d[a] = d[b];
				d[b] = td;
#pragma unroll
				for (int r = 0; r < N; r++) {
					const double tz = Z[r][a];
					Z[r][a] = Z[r][b];
					Z[r][b] = tz;
				}
			}
		}
#pragma unroll
	for (int c = 0; c < N; c++) {
		if (c < k) {
			s[c] = sqrt(fmax(d[c], 0.0));
#pragma unroll
			for (int a = 0; a < N; a++)
				if (a < k) U[a * k + c] = Z[a][c];
		}
	}
}

// SaiModel::operationalSpaceMatrices(J) for J (r x n): Lambda (r x r), N (n x n); Jbar is not needed by the callers
static __device__ __noinline__ void op_space(const double* J, int r, int n, const double* Minv, double* Lambda, double* Nout) {
	double MJt[MAXD * MAXD], A[MAXD * MAXD], Jbar[MAXD * MAXD];
	mm_bt(Minv, n, n, J, r, MJt);	 // n x r : Minv J^T
	mm(J, r, n, MJt, r, A);			 // r x r
	spd_inverse(A, r, Lambda);
	mm(MJt, n, r, Lambda, r, Jbar);	 // n x r
	mm(Jbar, n, r, J, n, Nout);		 // n x n
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++) Nout[i * n + j] = ((i == j) ? 1.0 : 0.0) - Nout[i * n + j];
}

// (J Mbinv J^T)^-1
static __device__ __noinline__ void lambda_of(const double* J, int r, int n, const double* Mbinv, double* out) {
	double MJt[MAXD * MAXD], A[MAXD * MAXD];
	mm_bt(Mbinv, n, n, J, r, MJt);
	mm(J, r, n, MJt, r, A);
	spd_inverse(A, r, out);
}

}  // namespace sg

// One full control cycle of ONE robot for an arbitrary task hierarchy (P.tasks), following the reference's
// RobotController loop (src/RobotController.cpp:68-118) with explicit N_prec chaining.  The task models are pure
// functions of the state, so model update and torque of task k are evaluated together, in hierarchy order.
template <int N>
static __device__ __noinline__ void generic_cycle_one(const OscProgram& P, const int64_t i, uint32_t status) {
	using namespace sg;
	const int64_t NR = P.n_robots;
	const DevModel& mdl = P.model;
	constexpr int n = N;

	double q[N], dq[N];
	for (int j = 0; j < N; j++) {
		q[j] = P.q[(int64_t)j * NR + i];
		dq[j] = P.dq[(int64_t)j * NR + i];
	}
	KinDyn<N> kd;
	forward_kinematics<N>(mdl, q, kd);
	mass_matrix<N, true>(mdl, kd);
	double M[N * N], Minv[N * N];
	for (int a = 0; a < N; a++)
		for (int b = 0; b < N; b++) M[a * N + b] = kd.M[a][b];
	spd_inverse(M, n, Minv);

	double Nprec[N * N];
	for (int a = 0; a < N * N; a++) Nprec[a] = ((a / N) == (a % N)) ? 1.0 : 0.0;
	double tau[N];
	for (int j = 0; j < N; j++) tau[j] = 0.0;

	for (int task = 0; task < P.n_tasks; task++) {
		double Nmat[N * N];	 // null space of this task (getTaskNullspace)
		if (P.tasks[task].type == OSC_TASK_MOTION_FORCE) {
			const DevMft& t = P.mft[P.tasks[task].index];
			const osc_mft_params& p = t.p;
			double* st = t.st;
			int32_t* ist = t.ist;
			const int r = t.rank;
			// _jacobian = P * JWorldFrame, _projected_jacobian = _jacobian * N_prec (MotionForceTask.cpp:261-264)
			double x[3], Rc[9];
			frame_pose<N>(kd, t.body, t.ctrl_R, t.ctrl_t, x, Rc);
			double JT0[N][6];
			point_jacobian_t<N>(mdl, kd, t.body, x, JT0);
			double J0[6 * N], J[6 * N];
			for (int j = 0; j < N; j++) {
				double v[3] = {JT0[j][0], JT0[j][1], JT0[j][2]}, w[3] = {JT0[j][3], JT0[j][4], JT0[j][5]};
				if (!t.full) {
					double tv[3], tw[3];
					mat3_vec(t.Pt, v, tv);
					mat3_vec(t.Pr, w, tw);
					for (int k = 0; k < 3; k++) {
						v[k] = tv[k];
						w[k] = tw[k];
					}
				}
				for (int k = 0; k < 3; k++) {
					J0[k * N + j] = v[k];
					J0[(3 + k) * N + j] = w[k];
				}
			}
			mm(J0, 6, n, Nprec, n, J);

			// ---- SingularityHandler::updateTaskModel (:75-228)
			constexpr int K = (N < 6) ? N : 6;	// thin SVD width
			double U[6 * K], sv[K], V[N * K];
			svd_rows6(J, n, U, sv, V);
			int n_ns = 0, n_s = 0;	// columns of the non-singular / singular task range
			double alpha = 1.0;
			if (sv[0] < p.s_abs_tol) {	// fully singular (:83-98)
				alpha = 0.0;
				n_ns = 0;
				n_s = r;
			} else if (r == 1) {
				n_ns = 1;
			} else {
				n_ns = r;
				for (int c = 1; c < r; c++) {
					const double icn = sv[c] / sv[0];
					if (icn < p.s_max) {
						alpha = fmin(fmax((icn - p.s_min) / (p.s_max - p.s_min), 0.0), 1.0);
						n_ns = c;
						n_s = r - c;
						break;
					}
				}
			}
			if (n_s > 0) status |= OSC_STATUS_SINGULAR_PATH;
			double Uns[6 * 6], Us[6 * 6], Vs[N * 6], Jns[6 * N], Js[6 * N];
			for (int a = 0; a < 6; a++) {
				for (int c = 0; c < n_ns; c++) Uns[a * n_ns + c] = U[a * K + c];
				for (int c = 0; c < n_s; c++) Us[a * n_s + c] = U[a * K + n_ns + c];
			}
			for (int a = 0; a < N; a++)
				for (int c = 0; c < n_s; c++) Vs[a * n_s + c] = V[a * K + n_ns + c];
			double Lns[36], Nns[N * N], Ls[36], Ljs[36], Jpost[6 * N];
			for (int a = 0; a < N * N; a++) Nns[a] = ((a / N) == (a % N)) ? 1.0 : 0.0;
			if (n_ns > 0) {
				mm_at(Uns, 6, n_ns, J, n, Jns);
				op_space(Jns, n_ns, n, Minv, Lns, Nns);
			}
			if (n_s > 0 && n_ns > 0) {
				mm_at(Us, 6, n_s, J, n, Js);
				lambda_of(Js, n_s, n, Minv, Ls);
			}
			const bool handling = p.singularity_handling_enabled != 0;
			bool have_post = false;
			if (n_s == 0 || !handling) {
				for (int a = 0; a < N * N; a++) Nmat[a] = Nns[a];
			} else if (n_ns == 0) {
				for (int a = 0; a < N * N; a++) Nmat[a] = Nprec[a];	 // fully singular: pass the task through (:149-151)
			} else {
				double Njs[N * N], T1[6 * N];
				mm_at(Vs, n, n_s, Nns, n, T1);	   // V_s^T N_ns
				mm(T1, n_s, n, Nprec, n, Jpost);  // ... N_prec
				op_space(Jpost, n_s, n, Minv, Ljs, Njs);
				mm(Njs, n, n, Nns, n, Nmat);
				have_post = true;
			}
			// decoupling (:160-225)
			double Lns_mod[36], Ls_mod[36], Ljs_mod[36];
			if (p.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES) {
				double Mb[N * N], Mbinv[N * N];
				for (int a = 0; a < N * N; a++) Mb[a] = M[a];
				for (int a = 0; a < N; a++)
					if (Mb[a * N + a] < p.bie_threshold) Mb[a * N + a] = p.bie_threshold;
				spd_inverse(Mb, n, Mbinv);
				if (n_ns > 0) lambda_of(Jns, n_ns, n, Mbinv, Lns_mod);
				if (n_s > 0 && n_ns > 0) lambda_of(Js, n_s, n, Mbinv, Ls_mod);
				if (have_post) lambda_of(Jpost, n_s, n, Mbinv, Ljs_mod);
			} else if (p.dynamic_decoupling_type == OSC_FULL_DYNAMIC_DECOUPLING) {
				for (int a = 0; a < n_ns * n_ns; a++) Lns_mod[a] = Lns[a];
				for (int a = 0; a < n_s * n_s; a++) Ls_mod[a] = Ls[a];
				for (int a = 0; a < n_s * n_s; a++) Ljs_mod[a] = Ljs[a];
			} else {
				for (int a = 0; a < 36; a++) {
					Lns_mod[a] = 0.0;
					Ls_mod[a] = 0.0;
					Ljs_mod[a] = 0.0;
				}
				for (int a = 0; a < n_ns; a++) Lns_mod[a * n_ns + a] = 1.0;
				for (int a = 0; a < n_s; a++) {
					Ls_mod[a * n_s + a] = 1.0;
					Ljs_mod[a * n_s + a] = 1.0;
				}
			}

			// ---- classifySingularity (:230-295)
			int32_t c1 = ist[(int64_t)MI_T1_COUNTER * NR + i], c2 = ist[(int64_t)MI_T2_COUNTER * NR + i];
			const int32_t n_types_prev = ist[(int64_t)MI_N_TYPES * NR + i];
			int32_t hist_head = ist[(int64_t)MI_HIST_HEAD * NR + i], hist_size = ist[(int64_t)MI_HIST_SIZE * NR + i];
			double q_prior[N];
			const bool upd = P.update_models != 0;
			if (upd && (n_types_prev == 0 || c2 > c1)) {
				for (int j = 0; j < N; j++) {
					q_prior[j] = q[j];
					st[(int64_t)(MC_Q_PRIOR + j) * NR + i] = q[j];
					st[(int64_t)(MC_DQ_PRIOR + j) * NR + i] = dq[j];
				}
			} else {
				for (int j = 0; j < N; j++) q_prior[j] = st[(int64_t)(MC_Q_PRIOR + j) * NR + i];
			}
			int n_types = 0;
			bool any_type1 = false;
			if (n_s == 0) {
				if (upd) {
					c1 = c2 = 0;
					hist_head = hist_size = 0;
				}
			} else {
				n_types = n_s;
				for (int c = 0; c < n_s; c++) {
					double qq[N];
					for (int j = 0; j < N; j++) qq[j] = q[j] + p.perturb_step_size * Vs[j * n_s + c];
					KinDyn<N> kp;
					forward_kinematics<N>(mdl, qq, kp);
					double xp[3], Rp[9], dphi[3];
					frame_pose<N>(kp, t.body, t.ctrl_R, t.ctrl_t, xp, Rp);
					orientation_error(Rp, Rc, dphi);
					double mot = 0.0;
					for (int k = 0; k < 3; k++) mot += (xp[k] - x[k]) * Us[k * n_s + c] + dphi[k] * Us[(3 + k) * n_s + c];
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

			// ---- MotionForceTask::computeTorques (:278-509) + SingularityHandler::computeTorques (:297-368)
			double fstar[6], F[6];
			double vv[3] = {0, 0, 0}, ww[3] = {0, 0, 0};
			for (int j = 0; j < N; j++)
				for (int k = 0; k < 3; k++) {
					vv[k] += JT0[j][k] * dq[j];
					ww[k] += JT0[j][3 + k] * dq[j];
				}
			mft_control_law(t, NR, i, x, Rc, vv, ww, P.write_observers != 0, fstar, F, status);
			double tt[N];
			for (int j = 0; j < N; j++) tt[j] = 0.0;
			double a[6], b[6], c[6];
			if (n_ns > 0) {
				mv_t(Uns, 6, n_ns, fstar, a);  // U_ns^T f*
				mv_t(Uns, 6, n_ns, F, b);
				const bool plain = (n_types != 0) && (p.dynamic_decoupling_type == OSC_IMPEDANCE);
				if (plain) {
					for (int k = 0; k < n_ns; k++) c[k] = a[k] + b[k];
				} else {
					mv(Lns_mod, n_ns, n_ns, a, c);
					for (int k = 0; k < n_ns; k++) c[k] += b[k];
				}
				mv_t(Jns, n_ns, n, c, tt);	// tau_ns
			}
			const bool blended = (n_types != 0) && (p.dynamic_decoupling_type != OSC_IMPEDANCE) && n_ns > 0 && handling && n_s > 0;
			if (blended) {
				double js[N], u[N], vt[6], tmp[6];
				if (c1 > c2 || p.enforce_type_1_strategy) {
					status |= OSC_STATUS_TYPE1;
					for (int j = 0; j < N; j++) u[j] = -p.kp_type_1 * (q[j] - q_prior[j]) - p.kv_type_1 * dq[j];
					mv_t(Vs, n, n_s, u, vt);
					mv(Ljs_mod, n_s, n_s, vt, tmp);
					mv_t(Jpost, n_s, n, tmp, js);
				} else {
					status |= OSC_STATUS_TYPE2;
					double dir[N];
					for (int j = 0; j < N; j++) {
						dir[j] = st[(int64_t)(MC_TYPE2_DIR + j) * NR + i];
						if (Vs[j * n_s + 0] != 0.0) {
							if (fabs(q[j] - mdl.q_upper[j]) < p.type_2_angle_threshold)
								dir[j] = -1.0;
							else if (fabs(q[j] - mdl.q_lower[j]) < p.type_2_angle_threshold)
								dir[j] = 1.0;
						}
						st[(int64_t)(MC_TYPE2_DIR + j) * NR + i] = dir[j];
					}
					double ff[6], nf = 0.0, fTd = 0.0;
					for (int k = 0; k < 6; k++) {
						ff[k] = fstar[k] + F[k];
						nf += ff[k] * ff[k];
					}
					nf = sqrt(nf);
					for (int k = 0; k < 6; k++) fTd += (nf > 0.0 ? ff[k] / nf : ff[k]) * Us[k * n_s + 0];
					for (int j = 0; j < N; j++) u[j] = dir[j] * fabs(fTd) * p.type_2_torque_ratio * mdl.effort[j];
					double js2[N];
					mv_t(Vs, n, n_s, u, vt);
					mv_t(Jpost, n_s, n, vt, js);
					for (int j = 0; j < N; j++) u[j] = -p.kv_type_2 * dq[j];
					mv_t(Vs, n, n_s, u, vt);
					mv(Ljs_mod, n_s, n_s, vt, tmp);
					mv_t(Jpost, n_s, n, tmp, js2);
					for (int j = 0; j < N; j++) js[j] += js2[j];
				}
				double ts[N];
				mv_t(Us, 6, n_s, fstar, a);
				mv_t(Us, 6, n_s, F, b);
				mv(Ls_mod, n_s, n_s, a, c);
				for (int k = 0; k < n_s; k++) c[k] += b[k];
				mv_t(Js, n_s, n, c, ts);
				for (int j = 0; j < N; j++) {
					if (isnan(ts[j])) {
						ts[j] = 0.0;
						status |= OSC_STATUS_NAN_SCRUBBED;
					} else if (ts[j] > mdl.effort[j])
						ts[j] = mdl.effort[j];
					else if (ts[j] < -mdl.effort[j])
						ts[j] = -mdl.effort[j];
					tt[j] += alpha * ts[j] + (1.0 - alpha) * js[j];
				}
			}
			// computeTorques(tau_prec) subtracts a term that is identically zero (MotionForceTask::_Lambda == 0, Appendix C1)
			for (int j = 0; j < N; j++) tau[j] += tt[j];
		} else {
			// ---- JointTask::updateTaskModel / computeTorques with a general N_prec and selection S (JointTask.cpp:218-356)
			const DevJt& jt = P.jt[P.tasks[task].index];
			const osc_joint_params& jp = jt.p;
			const int k = jt.k;
			double S[N * N], Jp[N * N];
			for (int a = 0; a < k; a++)
				for (int j = 0; j < N; j++) S[a * N + j] = jt.S[a][j];
			mm(S, k, n, Nprec, n, Jp);	// _projected_jacobian (k x n)
			// range basis (SaiModel::matrixRangeBasis, tolerance 1e-3)
			double Ur[N * N], sr[N];
			int kr = 0;
			if (sqrt(fro2(Jp, k * n)) >= 1e-3) {
				left_svd_gram<N>(Jp, k, n, Ur, sr);	 // Ur: k x k
				if (sr[0] >= 1e-3) {
					kr = k;
					for (int c = k - 1; c > 0; c--) {
						if (sr[c] / sr[0] < 1e-3)
							kr--;
						else
							break;
					}
				}
			}
			if (kr == 0) {
				status |= OSC_STATUS_ZERO_RANGE;
				for (int a = 0; a < N * N; a++) Nmat[a] = ((a / N) == (a % N)) ? 1.0 : 0.0;
			} else {
				double Ub[N * N];  // k x kr, identity when the range is the whole task space
				for (int a = 0; a < k; a++)
					for (int c = 0; c < kr; c++) Ub[a * kr + c] = (kr == k) ? ((a == c) ? 1.0 : 0.0) : Ur[a * k + c];
				double Jr[N * N], Mp[N * N], Mmod[N * N];
				mm_at(Ub, k, kr, Jp, n, Jr);  // U^T J_proj  (kr x n)
				op_space(Jr, kr, n, Minv, Mp, Nmat);
				if (jp.dynamic_decoupling_type == OSC_FULL_DYNAMIC_DECOUPLING) {
					for (int a = 0; a < kr * kr; a++) Mmod[a] = Mp[a];
				} else if (jp.dynamic_decoupling_type == OSC_BOUNDED_INERTIA_ESTIMATES) {
					double Mb[N * N], Mbi[N * N];
					for (int a = 0; a < N * N; a++) Mb[a] = M[a];
					for (int a = 0; a < N; a++)
						if (Mb[a * N + a] < jp.bie_threshold) Mb[a * N + a] = jp.bie_threshold;
					spd_inverse(Mb, n, Mbi);
					lambda_of(Jr, kr, n, Mbi, Mmod);
				} else {
					for (int a = 0; a < kr * kr; a++) Mmod[a] = ((a / kr) == (a % kr)) ? 1.0 : 0.0;
				}
				double pid[N], acc[N];
				joint_control_law_rt<N>(jt, NR, i, q, dq, k, pid, acc);
				double a1[N], a2[N], f[N], g[N];
				mv_t(Ub, k, kr, acc, a1);
				mv(Mp, kr, kr, a1, f);
				mv_t(Ub, k, kr, pid, a2);
				mv(Mmod, kr, kr, a2, g);
				for (int c = 0; c < kr; c++) f[c] += g[c];
				if (P.use_prev_torques) {  // - J_proj^T U M_partial U^T S Minv tau_prec
					double mt[N], smt[N], ut[N], w[N];
					mv(Minv, n, n, tau, mt);
					mv(S, k, n, mt, smt);
					mv_t(Ub, k, kr, smt, ut);
					mv(Mp, kr, kr, ut, w);
					for (int c = 0; c < kr; c++) f[c] -= w[c];
				}
				double uf[N], tj[N];
				mv(Ub, k, kr, f, uf);
				mv_t(Jp, k, n, uf, tj);
				for (int j = 0; j < N; j++) tau[j] += tj[j];
			}
		}
		// N_prec <- task.N * N_prec  (getTaskAndPreviousNullspace, RobotController.cpp:75)
		double Nn[N * N];
		mm(Nmat, n, n, Nprec, n, Nn);
		for (int a = 0; a < N * N; a++) Nprec[a] = Nn[a];
	}

	if (P.torque_saturation)
		for (int j = 0; j < N; j++) tau[j] = fmin(fmax(tau[j], -mdl.effort[j]), mdl.effort[j]);
	if (P.gravity_comp)
		for (int j = 0; j < N; j++) tau[j] += kd.g[j];
	for (int j = 0; j < N; j++) P.tau[(int64_t)j * NR + i] = tau[j];
	P.status[i] = status;
}

// SVD-path kernel: the robots in `sing_list` (count on the device, written by the fast kernel of the same cycle).
// Fixed grid, grid-stride over the list: with an empty list every thread exits at once.
template <int N>
__global__ void __launch_bounds__(64, OSC_GENERIC_MIN_BLOCKS) osc_singular_kernel(const __grid_constant__ OscProgram P) {
	// programmatic dependent launch: wait until the fast kernel of this cycle has completed and flushed its writes
	asm volatile("griddepcontrol.launch_dependents;");
	asm volatile("griddepcontrol.wait;" ::: "memory");
	const int32_t count = P.sing_count[P.sing_parity];
	const int stride = gridDim.x * blockDim.x;
	for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < count; slot += stride)
		generic_cycle_one<N>(P, (int64_t)P.sing_list[slot], OSC_STATUS_SINGULAR_PATH);
	publish_general_done(P, count);
}

// Whole-batch kernel for hierarchies without a specialised fast kernel (partial joint tasks, several motion-force
// tasks, ...): every robot takes the general path.
template <int N>
__global__ void __launch_bounds__(64, OSC_GENERIC_MIN_BLOCKS) osc_generic_kernel(const __grid_constant__ OscProgram P) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P.n_robots) return;
	generic_cycle_one<N>(P, i, 0u);
}

}  // namespace osc
