// Small fixed-size FP64 linear algebra for one-robot-per-thread kernels.
// Every loop bound is a template constant so that, after full unrolling, all array
// indices are compile-time constants and the arrays live in registers (or spill to
// thread-local memory under register pressure -- never dynamic indexing).
#pragma once
#include <cuda_runtime.h>

#define DEVI __device__ __forceinline__

namespace osc {

DEVI void cross3(const double a[3], const double b[3], double o[3]) {
	o[0] = a[1] * b[2] - a[2] * b[1];
	o[1] = a[2] * b[0] - a[0] * b[2];
	o[2] = a[0] * b[1] - a[1] * b[0];
}
DEVI double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// o = A(3x3 row-major) * v
DEVI void mat3_vec(const double A[9], const double v[3], double o[3]) {
	o[0] = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
	o[1] = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
	o[2] = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
}
// o = A^T * v
DEVI void mat3t_vec(const double A[9], const double v[3], double o[3]) {
	o[0] = A[0] * v[0] + A[3] * v[1] + A[6] * v[2];
	o[1] = A[1] * v[0] + A[4] * v[1] + A[7] * v[2];
	o[2] = A[2] * v[0] + A[5] * v[1] + A[8] * v[2];
}
// C = A * B (3x3 row-major)
DEVI void mat3_mul(const double A[9], const double B[9], double C[9]) {
#pragma unroll
	for (int i = 0; i < 3; i++)
#pragma unroll
		for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// C = A * B^T
DEVI void mat3_mul_bt(const double A[9], const double B[9], double C[9]) {
#pragma unroll
	for (int i = 0; i < 3; i++)
#pragma unroll
		for (int j = 0; j < 3; j++)
			C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}

#ifndef OSC_FP32_BUILD
// ---- lean FP64 reciprocal / reciprocal square root / sine-cosine --------------------------------------------------
// The CUDA library versions carry special-case branches (subnormals, infinities, huge arguments) that cost a
// predicate test, a convergence barrier and a cold subroutine per call site: about 35 instructions for rsqrt(), 80 for
// sincos(), inlined dozens of times in the fused kernel.  The arguments here are pivots of positive definite matrices,
// squared column norms and joint angles, so the hardware seed (MUFU.RSQ64H / MUFU.RCP64H, about 20 bits) plus
// straight-line Newton steps is enough: results are accurate to 1-2 ulp for normal positive arguments.
DEVI double rsqrt_pos(double d) {
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
	double h = d * y;
	double e = fma(-h, y, 1.0);	 // 1 - d y^2
	y = fma(y * e, fma(e, 0.375, 0.5), y);	// third order: y (1 + e/2 + 3 e^2/8)
	h = d * y;
	e = fma(-h, y, 1.0);
	return fma(y * e, 0.5, y);
}
DEVI double rcp_nz(double d) {
	double y;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
	double e = fma(-d, y, 1.0);
	y = fma(y, e, y);
	e = fma(-d, y, 1.0);
	y = fma(y, e, y);
	e = fma(-d, y, 1.0);
	return fma(y, e, y);
}
// sqrt of a positive normal number
DEVI double sqrt_pos(double d) {
	const double y = rsqrt_pos(d);
	const double s = d * y;
	return fma(fma(-s, s, d), 0.5 * y, s);
}
// a / b for finite a and normal b
DEVI double div_nz(double a, double b) {
	const double y = rcp_nz(b);
	const double q = a * y;
	return fma(fma(-b, q, a), y, q);
}
// sin and cos of a joint angle.  Cody-Waite reduction by pi/2 in two pieces (exact enough for |x| < 1e4: the error
// of the reduced argument stays below 1e-19 |k|) and the minimax polynomials of fdlibm's __kernel_sin/__kernel_cos
// on [-pi/4, pi/4]; larger arguments take the library routine.
static __device__ __noinline__ void sincos_far(double x, double* sn, double* cs) { sincos(x, sn, cs); }
// polynomial coefficients as constant-bank operands (a 64-bit immediate costs two moves per use)
__constant__ double kSinCoef[6] = {1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
									-1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01};
__constant__ double kCosCoef[6] = {-1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
									2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};
__constant__ double kPio2[3] = {0.63661977236758134308, 1.57079632679489655800e+00, 6.12323399573676603587e-17};
DEVI void sincos_joint(double x, double* sn, double* cs) {
	if (!(fabs(x) < 1.0e4)) {
		sincos_far(x, sn, cs);
		return;
	}
	const int q = __double2int_rn(x * kPio2[0]);
	const double k = (double)q;
	double r = fma(-k, kPio2[1], x);
	r = fma(-k, kPio2[2], r);
	const double z = r * r;
	double ps = fma(z, kSinCoef[0], kSinCoef[1]);
	ps = fma(z, ps, kSinCoef[2]);
	ps = fma(z, ps, kSinCoef[3]);
	ps = fma(z, ps, kSinCoef[4]);
	ps = fma(z, ps, kSinCoef[5]);
	const double s = fma(z * r, ps, r);
	double pc = fma(z, kCosCoef[0], kCosCoef[1]);
	pc = fma(z, pc, kCosCoef[2]);
	pc = fma(z, pc, kCosCoef[3]);
	pc = fma(z, pc, kCosCoef[4]);
	pc = fma(z, pc, kCosCoef[5]);
	const double c = fma(z * z, pc, fma(z, -0.5, 1.0));
	const double a = (q & 1) ? c : s;
	const double b = (q & 1) ? s : c;
	*sn = (q & 2) ? -a : a;
	*cs = ((q + 1) & 2) ? -b : b;
}

// ---- asynchronous global -> shared copies of single doubles (LDGSTS): the data lands in this thread's shared-memory
// slots without passing through registers, so a load can be issued thousands of cycles before its use
DEVI void cp_async8(double* smem_dst, const double* gsrc) {
	const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
DEVI void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#else
// the same routines for the single-precision instantiation (gen_fp32.py copies this region verbatim): hardware seed plus
// one Newton step (about 1 ulp), the library sine-cosine, and a plain converting load in place of the 8-byte cp.async
DEVI float rsqrt_pos(float d) {
	float y;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
	const float e = fmaf(-d * y, y, 1.0f);
	return fmaf(y * e, 0.5f, y);
}
DEVI float rcp_nz(float d) {
	float y;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
	const float e = fmaf(-d, y, 1.0f);
	return fmaf(y, e, y);
}
DEVI float sqrt_pos(float d) {
	const float y = rsqrt_pos(d);
	const float s = d * y;
	return fmaf(fmaf(-s, s, d), 0.5f * y, s);
}
DEVI float div_nz(float a, float b) {
	const float y = rcp_nz(b);
	const float q = a * y;
	return fmaf(fmaf(-b, q, a), y, q);
}
DEVI void sincos_joint(float x, float* sn, float* cs) { sincosf(x, sn, cs); }
DEVI void cp_async8(float* smem_dst, const gdouble* gsrc) { *smem_dst = (float)*gsrc; }
DEVI void cp_async_wait_all() {}
#endif  // OSC_FP32_BUILD

// In-place Cholesky of the lower triangle of a symmetric positive definite matrix: A = L L^T.
// invd[j] = 1 / L[j][j] is kept so that the triangular solves multiply instead of divide (one FP64 reciprocal
// per pivot instead of one division per solve step).  Returns false when a pivot is not positive.
template <int N>
DEVI bool cholesky_lower(double (&A)[N][N], double (&invd)[N]) {
	bool ok = true;
#pragma unroll
	for (int j = 0; j < N; j++) {
		double d = A[j][j];
#pragma unroll
		for (int k = 0; k < j; k++) d -= A[j][k] * A[j][k];
		ok = ok && (d > 0.0);
		const double inv = rsqrt_pos(d);
		invd[j] = inv;
		A[j][j] = d * inv;
#pragma unroll
		for (int i = j + 1; i < N; i++) {
			double s = A[i][j];
#pragma unroll
			for (int k = 0; k < j; k++) s -= A[i][k] * A[j][k];
			A[i][j] = s * inv;
		}
	}
	return ok;
}

// x <- L^-1 x   (L lower triangular, N x N)
template <int N, class LT>
DEVI void solve_lower(const LT& L, const double (&invd)[N], double (&x)[N]) {
#pragma unroll
	for (int i = 0; i < N; i++) {
		double s = x[i];
#pragma unroll
		for (int k = 0; k < i; k++) s -= L[i][k] * x[k];
		x[i] = s * invd[i];
	}
}
// x <- L^-T x
template <int N, class LT>
DEVI void solve_lower_t(const LT& L, const double (&invd)[N], double (&x)[N]) {
#pragma unroll
	for (int i = N - 1; i >= 0; i--) {
		double s = x[i];
#pragma unroll
		for (int k = i + 1; k < N; k++) s -= L[k][i] * x[k];
		x[i] = s * invd[i];
	}
}
// y = L x
template <int N, class LT>
DEVI void mul_lower(const LT& L, const double (&x)[N], double (&y)[N]) {
#pragma unroll
	for (int i = 0; i < N; i++) {
		double s = 0.0;
#pragma unroll
		for (int k = 0; k <= i; k++) s += L[i][k] * x[k];
		y[i] = s;
	}
}
// x <- (L L^T)^-1 x
template <int N, class LT>
DEVI void solve_spd(const LT& L, const double (&invd)[N], double (&x)[N]) {
	solve_lower<N>(L, invd, x);
	solve_lower_t<N>(L, invd, x);
}

// x <- (R^T R)^-1 x  with R upper triangular R x R stored in X[C+i][j], i <= j, rinv[i] = 1 / R[i][i]
template <int N, int R, int C, class XT>
DEVI void solve_rtr(const XT& X, const double (&rinv)[R], double (&x)[R]) {
#pragma unroll
	for (int i = 0; i < R; i++) {  // R^T z = x (forward)
		double s = x[i];
#pragma unroll
		for (int k = 0; k < i; k++) s -= X[C + k][i] * x[k];
		x[i] = s * rinv[i];
	}
#pragma unroll
	for (int i = R - 1; i >= 0; i--) {	// R y = z (backward)
		double s = x[i];
#pragma unroll
		for (int k = i + 1; k < R; k++) s -= X[C + i][k] * x[k];
		x[i] = s * rinv[i];
	}
}

// Householder QR of rows C..N-1 of X (N x R), in place:
//   on exit X[C+i][j] (i <= j) holds R; reflector j is v_j with v_j[C+j] = vhead[j] and
//   v_j[C+j+1..N-1] stored in X below the diagonal; H_j = I - beta[j] v_j v_j^T;  rinv[j] = 1 / R[j][j].
template <int N, int R, int C, class XT>
DEVI void householder_qr(XT& X, double (&vhead)[R], double (&beta)[R], double (&rinv)[R]) {
#pragma unroll
	for (int j = 0; j < R; j++) {
		constexpr int dummy = 0;
		(void)dummy;
		const int k = C + j;
		double nrm2 = 0.0;
#pragma unroll
		for (int i = k; i < N; i++) nrm2 += X[i][j] * X[i][j];
		const double inrm = rsqrt_pos(nrm2);
		const double nrm = nrm2 * inrm;
		const double x0 = X[k][j];
		const double alpha = (x0 >= 0.0) ? -nrm : nrm;
		rinv[j] = (x0 >= 0.0) ? -inrm : inrm;
		const double v0 = x0 - alpha;
		// v^T v = nrm2 - 2 alpha x0 + alpha^2 = 2 (nrm2 - alpha x0)
		const double vv = 2.0 * (nrm2 - alpha * x0);
		// beta = 2 / v^T v = 1 / (nrm (nrm + |x0|)):  one reciprocal, no division
		const double b = (vv > 0.0) ? 2.0 * rcp_nz(vv) : 0.0;
		vhead[j] = v0;
		beta[j] = b;
#pragma unroll
		for (int jj = j + 1; jj < R; jj++) {
			double s = v0 * X[k][jj];
#pragma unroll
			for (int i = k + 1; i < N; i++) s += X[i][j] * X[i][jj];
			s *= b;
			X[k][jj] -= s * v0;
#pragma unroll
			for (int i = k + 1; i < N; i++) X[i][jj] -= s * X[i][j];
		}
		X[k][j] = alpha;
	}
}

// x <- H_1 ... H_R x   (reflectors from householder_qr, acting on rows C..N-1)
template <int N, int R, int C, class XT>
DEVI void apply_q(const XT& X, const double (&vhead)[R], const double (&beta)[R], double (&x)[N]) {
#pragma unroll
	for (int j = R - 1; j >= 0; j--) {
		const int k = C + j;
		double s = vhead[j] * x[k];
#pragma unroll
		for (int i = k + 1; i < N; i++) s += X[i][j] * x[i];
		s *= beta[j];
		x[k] -= s * vhead[j];
#pragma unroll
		for (int i = k + 1; i < N; i++) x[i] -= s * X[i][j];
	}
}
// x <- H_R ... H_1 x
template <int N, int R, int C, class XT>
DEVI void apply_qt(const XT& X, const double (&vhead)[R], const double (&beta)[R], double (&x)[N]) {
#pragma unroll
	for (int j = 0; j < R; j++) {
		const int k = C + j;
		double s = vhead[j] * x[k];
#pragma unroll
		for (int i = k + 1; i < N; i++) s += X[i][j] * x[i];
		s *= beta[j];
		x[k] -= s * vhead[j];
#pragma unroll
		for (int i = k + 1; i < N; i++) x[i] -= s * X[i][j];
	}
}

// Sufficient test for  s_min(J)/s_max(J) >= thr  on the R rows of J (stored as JT: N x R):
// with G = J J^T,  lambda_max(G) <= (tr G^8)^(1/8) =: hi;  G - thr^2 hi I positive definite
// implies lambda_min >= thr^2 hi >= thr^2 lambda_max.  Never accepts a singular robot;
// rejects a thin band (< 1%) of non-singular ones, which then take the SVD path.
template <int N, int R>
DEVI bool sound_nonsingular(const double (&JT)[N][R], double thr, double abs_tol) {
	double G[R][R];
#pragma unroll
	for (int a = 0; a < R; a++)
#pragma unroll
		for (int b = 0; b <= a; b++) {
			double s = 0.0;
#pragma unroll
			for (int i = 0; i < N; i++) s += JT[i][a] * JT[i][b];
			G[a][b] = s;
			G[b][a] = s;
		}
	double tr = 0.0;
#pragma unroll
	for (int a = 0; a < R; a++) tr += G[a][a];
	// sigma_0 >= sqrt(tr/R); require it comfortably above the absolute tolerance (SingularityHandler.cpp:83)
	if (!(tr > (double)R * abs_tol * abs_tol)) return false;
	if (R == 1) return true;
	double G2[R][R];
#pragma unroll
	for (int a = 0; a < R; a++)
#pragma unroll
		for (int b = 0; b <= a; b++) {
			double s = 0.0;
#pragma unroll
			for (int i = 0; i < R; i++) s += G[a][i] * G[i][b];
			G2[a][b] = s;
			G2[b][a] = s;
		}
	double t8 = 0.0;  // tr(G^8) = ||G^4||_F^2,  G^4 = G2*G2
#pragma unroll
	for (int a = 0; a < R; a++)
#pragma unroll
		for (int b = 0; b <= a; b++) {
			double s = 0.0;
#pragma unroll
			for (int i = 0; i < R; i++) s += G2[a][i] * G2[i][b];
			t8 += (a == b) ? s * s : 2.0 * s * s;
		}
	const double hi = sqrt_pos(sqrt_pos(sqrt_pos(t8)));
	const double shift = thr * thr * hi;
#pragma unroll
	for (int a = 0; a < R; a++) G[a][a] -= shift;
	double invd[R];
	return cholesky_lower<R>(G, invd);
}

// Same sufficient test starting from the Gram matrix G = J J^T (full symmetric storage, destroyed).
template <int R>
DEVI bool sound_nonsingular_gram(double (&G)[R][R], double thr, double abs_tol) {
	double tr = 0.0;
#pragma unroll
	for (int a = 0; a < R; a++) tr += G[a][a];
	if (!(tr > (double)R * abs_tol * abs_tol)) return false;
	if (R == 1) return true;
	double G2[R][R];
#pragma unroll
	for (int a = 0; a < R; a++)
#pragma unroll
		for (int b = 0; b <= a; b++) {
			double s = 0.0;
#pragma unroll
			for (int i = 0; i < R; i++) s += G[a][i] * G[i][b];
			G2[a][b] = s;
			G2[b][a] = s;
		}
	double t8 = 0.0;  // tr(G^8) = ||G^4||_F^2,  G^4 = G2*G2
#pragma unroll
	for (int a = 0; a < R; a++)
#pragma unroll
		for (int b = 0; b <= a; b++) {
			double s = 0.0;
#pragma unroll
			for (int i = 0; i < R; i++) s += G2[a][i] * G2[i][b];
			t8 += (a == b) ? s * s : 2.0 * s * s;
		}
	const double hi = sqrt_pos(sqrt_pos(sqrt_pos(t8)));
	const double shift = thr * thr * hi;
#pragma unroll
	for (int a = 0; a < R; a++) G[a][a] -= shift;
	double invd[R];
	return cholesky_lower<R>(G, invd);
}

// ---- matrices of one thread, addressed M[r][c] by the templates above, either in registers or in shared memory
// (element e of thread t at b[e * STRIDE + t]: conflict-free, every address is base + constant after unrolling).
template <int ROWS, int COLS>
struct RegMat {
	double a[ROWS][COLS];
	DEVI double* operator[](int r) { return a[r]; }
	DEVI const double* operator[](int r) const { return a[r]; }
};
template <int COLS, int STRIDE>
struct SmMat {
	double* b;
	struct Row {
		double* p;
		DEVI double& operator[](int c) const { return p[c * STRIDE]; }
	};
	DEVI Row operator[](int r) const { return Row{b + r * COLS * STRIDE}; }
};
template <int STRIDE>
struct SmLowerMat {	 // packed lower triangle
	double* b;
	struct Row {
		double* p;
		DEVI double& operator[](int c) const { return p[c * STRIDE]; }
	};
	DEVI Row operator[](int r) const { return Row{b + (r * (r + 1) / 2) * STRIDE}; }
};

// ---- lower-triangular factor staged in shared memory: element (r, c) of this thread's matrix at
// b[(r (r + 1) / 2 + c) * STRIDE], the reciprocal diagonal behind it.  STRIDE is the (compile-time) block size, so
// every address is base + constant.
template <int N, int STRIDE>
struct SmTri {
	double* b;
	DEVI double L(int r, int c) const { return b[(r * (r + 1) / 2 + c) * STRIDE]; }
	DEVI double invd(int r) const { return b[(N * (N + 1) / 2 + r) * STRIDE]; }
	DEVI void store(const double (&A)[N][N], const double (&inv)[N]) {
#pragma unroll
		for (int r = 0; r < N; r++) {
#pragma unroll
			for (int c = 0; c <= r; c++) b[(r * (r + 1) / 2 + c) * STRIDE] = A[r][c];
			b[(N * (N + 1) / 2 + r) * STRIDE] = inv[r];
		}
	}
	// x <- L^-1 x
	DEVI void solve_lower(double (&x)[N]) const {
#pragma unroll
		for (int i = 0; i < N; i++) {
			double s = x[i];
#pragma unroll
			for (int k = 0; k < i; k++) s -= L(i, k) * x[k];
			x[i] = s * invd(i);
		}
	}
	// x <- L^-T x
	DEVI void solve_lower_t(double (&x)[N]) const {
#pragma unroll
		for (int i = N - 1; i >= 0; i--) {
			double s = x[i];
#pragma unroll
			for (int k = i + 1; k < N; k++) s -= L(k, i) * x[k];
			x[i] = s * invd(i);
		}
	}
	// y = L x
	DEVI void mul_lower(const double (&x)[N], double (&y)[N]) const {
#pragma unroll
		for (int i = 0; i < N; i++) {
			double s = 0.0;
#pragma unroll
			for (int k = 0; k <= i; k++) s += L(i, k) * x[k];
			y[i] = s;
		}
	}
};

}  // namespace osc
