// Host build of the device eigen-solver of the blending path (csrc/osc_eig6.h) for tests/test_eig6.py.
#include "../../sai_primitives_b200/csrc/osc_eig6.h"

extern "C" void eig6_probe(const double* G, int count, double* Z_out, double* d_out) {
	for (int n = 0; n < count; n++) {
		double A[6][6], Z[6][6], d[6];
		for (int a = 0; a < 6; a++)
			for (int b = 0; b < 6; b++) A[a][b] = G[n * 36 + a * 6 + b];
		osc::sym_eig6(A, Z, d);
		for (int a = 0; a < 6; a++) {
			d_out[n * 6 + a] = d[a];
			for (int b = 0; b < 6; b++) Z_out[n * 36 + a * 6 + b] = Z[a][b];
		}
	}
}

template <int M>
static void probe_m(const double* G, int count, double* Z_out, double* d_out) {
	for (int n = 0; n < count; n++) {
		double A[M][M], Z[M][M], d[M];
		for (int a = 0; a < M; a++)
			for (int b = 0; b < M; b++) A[a][b] = G[n * M * M + a * M + b];
		osc::sym_eig<M>(A, Z, d);
		for (int a = 0; a < M; a++) {
			d_out[n * M + a] = d[a];
			for (int b = 0; b < M; b++) Z_out[n * M * M + a * M + b] = Z[a][b];
		}
	}
}
extern "C" void eig_probe_m(int M, const double* G, int count, double* Z_out, double* d_out) {
	if (M == 4) probe_m<4>(G, count, Z_out, d_out);
	if (M == 7) probe_m<7>(G, count, Z_out, d_out);
	if (M == 8) probe_m<8>(G, count, Z_out, d_out);
}
