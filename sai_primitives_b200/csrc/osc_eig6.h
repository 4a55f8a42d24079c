// Eigen-decomposition of the symmetric 6 x 6 Gram matrix J J^T of the singular / blending path (osc_blend.cuh): its
// eigenvectors are the left singular vectors the reference takes from Eigen::JacobiSVD of the task Jacobian
// (SingularityHandler.cpp:75-120), its eigenvalues the squared singular values.
//
// Householder tridiagonalisation followed by implicit QL sweeps with Wilkinson shifts (the EISPACK tred2 / tql2 pair).
// About a quarter of the arithmetic of the cyclic Jacobi sweeps it replaces (which cost more than the kinematics,
// the mass matrix and the Jacobian together: profiles/r02_summary.md).  Every array index is a compile-time constant
// after unrolling, so the matrices stay in registers on the device; the file also compiles for the host
// (tests/cpp/eig6_host_probe.cpp checks it against LAPACK through numpy).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define EIG6_HD __host__ __device__ __forceinline__
#else
#define EIG6_HD inline
#endif

namespace osc {

// square root / reciprocal: the lean device versions of osc_math.cuh when they are in scope, the library ones otherwise
#if defined(__CUDA_ARCH__) && defined(OSC_EIG6_LEAN_MATH)
#define EIG6_SQRT(x) sqrt_pos(x)
#define EIG6_RCP(x) rcp_nz(x)
#else
#define EIG6_SQRT(x) sqrt(x)
#define EIG6_RCP(x) (1.0 / (x))
#endif

// A (symmetric M x M, both triangles filled; destroyed) = Z diag(d) Z^T, eigenvalues in no particular order.  M = 6 is the Gram
// matrix of the task Jacobian; the rolled general path also uses M = number of joints for the range basis of the joint task's
// projected Jacobian (osc_singular.cuh).
template <int M>
EIG6_HD void sym_eig(double (&A)[M][M], double (&Z)[M][M], double (&d)[M]) {
	constexpr double kTiny = 1e-290, kEps = 1.1102230246251565e-16;
	double e[M];
#pragma unroll
	for (int a = 0; a < M; a++)
#pragma unroll
		for (int b = 0; b < M; b++) Z[a][b] = (a == b) ? 1.0 : 0.0;
	// ---- tridiagonalisation: reflection k clears A[k+2.., k]
#pragma unroll
	for (int k = 0; k < M - 2; k++) {
		double sigma = 0.0;
#pragma unroll
		for (int r = k + 2; r < M; r++) sigma += A[r][k] * A[r][k];
		const double x0 = A[k + 1][k];
		if (sigma > kTiny) {
			const double nrm = EIG6_SQRT(x0 * x0 + sigma);
			const double alpha = (x0 >= 0.0) ? -nrm : nrm;
			double v[M];
#pragma unroll
			for (int r = 0; r < M; r++) v[r] = (r == k + 1) ? x0 - alpha : (r > k + 1 ? A[r][k] : 0.0);
			const double beta = 2.0 * EIG6_RCP(v[k + 1] * v[k + 1] + sigma);
			double pv[M], K = 0.0;
#pragma unroll
			for (int r = k + 1; r < M; r++) {
				double s = 0.0;
#pragma unroll
				for (int c = k + 1; c < M; c++) s += A[r][c] * v[c];
				pv[r] = beta * s;
				K += pv[r] * v[r];
			}
			K *= 0.5 * beta;
#pragma unroll
			for (int r = k + 1; r < M; r++) pv[r] -= K * v[r];	// w
#pragma unroll
			for (int r = k + 1; r < M; r++)
#pragma unroll
				for (int c = k + 1; c <= r; c++) {
					const double t = A[r][c] - v[r] * pv[c] - pv[r] * v[c];
					A[r][c] = t;
					A[c][r] = t;
				}
			A[k + 1][k] = alpha;
			// Z <- Z H
#pragma unroll
			for (int r = 0; r < M; r++) {
				double s = 0.0;
#pragma unroll
				for (int c = k + 1; c < M; c++) s += Z[r][c] * v[c];
				s *= beta;
#pragma unroll
				for (int c = k + 1; c < M; c++) Z[r][c] -= s * v[c];
			}
		}
	}
#pragma unroll
	for (int k = 0; k < M; k++) {
		d[k] = A[k][k];
		e[k] = (k < M - 1) ? A[k + 1 < M ? k + 1 : M - 1][k] : 0.0;
	}
	// ---- implicit QL: e[i] couples i and i + 1
#pragma unroll
	for (int l = 0; l < M - 1; l++) {
		for (int iter = 0; iter < 40; iter++) {
			int m = M - 1;
#pragma unroll
			for (int mm = M - 2; mm >= l; mm--)
				if (fabs(e[mm]) <= kEps * (fabs(d[mm]) + fabs(d[mm + 1]))) m = mm;
			if (m == l) break;
			double dm = d[M - 1];
#pragma unroll
			for (int mm = M - 2; mm > l; mm--)
				if (m == mm) dm = d[mm];
			double g = (d[l + 1] - d[l]) * (0.5 * EIG6_RCP(e[l]));
			double r = EIG6_SQRT(g * g + 1.0);
			g = dm - d[l] + e[l] * EIG6_RCP(g + (g >= 0.0 ? r : -r));
			double s = 1.0, c = 1.0, p = 0.0;
			bool under = false;
#pragma unroll
			for (int i = M - 2; i >= l; i--) {
				if (i < m && !under) {
					const double f = s * e[i], b = c * e[i];
					const double r2 = f * f + g * g;
					if (!(r2 > kTiny)) {  // the rotation is undefined: deflate here and start over
						e[i + 1] = 0.0;
						d[i + 1] -= p;
						under = true;
					} else {
						r = EIG6_SQRT(r2);
						e[i + 1] = r;
						const double ir = EIG6_RCP(r);
						s = f * ir;
						c = g * ir;
						g = d[i + 1] - p;
						r = (d[i] - g) * s + 2.0 * c * b;
						p = s * r;
						d[i + 1] = g + p;
						g = c * r - b;
#pragma unroll
						for (int k = 0; k < M; k++) {
							const double zk = Z[k][i + 1];
							Z[k][i + 1] = s * Z[k][i] + c * zk;
							Z[k][i] = c * Z[k][i] - s * zk;
						}
					}
				}
			}
			if (!under) {
				d[l] -= p;
				e[l] = g;
			}
#pragma unroll
			for (int mm = M - 1; mm > l; mm--)
				if (m == mm) e[mm] = 0.0;
		}
	}
}

EIG6_HD void sym_eig6(double (&A)[6][6], double (&Z)[6][6], double (&d)[6]) { sym_eig<6>(A, Z, d); }

}  // namespace osc
