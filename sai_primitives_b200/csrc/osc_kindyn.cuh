// Per-robot kinematics / dynamics stage (north_star subsystem (a)): forward kinematics,
// joint axes/origins for the 6 x n point Jacobian, composite-rigid-body mass matrix and
// gravity vector of a serial chain, all in world coordinates.  Replaces what the reference
// obtains from sai-model: updateModel / M / JWorldFrame / positionInWorld / rotationInWorld /
// jointGravityVector (call sites: SURVEY.md section 8c).
#pragma once
// Phase barriers of the fused cycle kernel: the warps of a block walk the long, fully unrolled instruction stream
// together, so an instruction-cache line fetched for one warp serves the others (measured +6 % on B200,
// profiles/).  Legal because no thread of the block leaves the kernel early.
#include "osc_dev_types.h"
#include "osc_math.cuh"
#if defined(OSC_TRACE)
// Tuning builds only (tools/trace_phases.py): thread 0 of every block stamps the SM clock and the global timer at each
// phase boundary into a buffer that the launcher dumps to $OSC_TRACE_FILE.
namespace osc {
__device__ unsigned long long* g_trace = nullptr;
__shared__ int s_trace_k;
DEVI void trace_point(bool sync) {
	if (sync) __syncthreads();
	if (threadIdx.x == 0 && g_trace) {
		const int k = s_trace_k++;
		unsigned long long g;
		unsigned smid;
		asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g));
		asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
		if (k < 31) {
			g_trace[((size_t)blockIdx.x * 32 + k) * 2] = (unsigned long long)clock64();
			g_trace[((size_t)blockIdx.x * 32 + k) * 2 + 1] = g;
		}
		g_trace[((size_t)blockIdx.x * 32 + 31) * 2] = smid;
		g_trace[((size_t)blockIdx.x * 32 + 31) * 2 + 1] = k + 1;
	}
}
}  // namespace osc
#ifndef OSC_NO_PHASE_SYNC
#define OSC_LS() osc::trace_point(true)
#else
#define OSC_LS() osc::trace_point(false)
#endif
#elif !defined(OSC_NO_PHASE_SYNC)
#define OSC_LS() __syncthreads()
#else
#define OSC_LS() ((void)0)
#endif

namespace osc {

template <int N>
struct KinDyn {
	double a[N][3];	 // joint axis, world
	double p[N][3];	 // joint origin (= body frame origin), world
	double Rb[N][9]; // body orientation, world
	double M[N][N];	 // joint-space mass matrix (full symmetric storage)
	double g[N];	 // jointGravityVector
};

// Rotation about a unit axis (Rodrigues); exact pattern for coordinate axes.
DEVI void axis_angle(const double ax[3], double s, double c, double R[9]) {
	const double v = 1.0 - c;
	const double x = ax[0], y = ax[1], z = ax[2];
	R[0] = c + x * x * v;
	R[1] = x * y * v - z * s;
	R[2] = x * z * v + y * s;
	R[3] = y * x * v + z * s;
	R[4] = c + y * y * v;
	R[5] = y * z * v - x * s;
	R[6] = z * x * v - y * s;
	R[7] = z * y * v + x * s;
	R[8] = c + z * z * v;
}

template <int N>
DEVI void forward_kinematics(const DevModel& m, const double (&q)[N], KinDyn<N>& kd) {
	double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	double p[3] = {0, 0, 0};
#pragma unroll
	for (int i = 0; i < N; i++) {
		double t[3];
		mat3_vec(R, m.t_fix[i], t);
		p[0] += t[0];
		p[1] += t[1];
		p[2] += t[2];
		double Rn[9];
		mat3_mul(R, m.R_fix[i], Rn);
		if (m.jtype[i] == 0) {
			double s, c;
			sincos(q[i], &s, &c);
			double Rq[9];
			axis_angle(m.axis[i], s, c, Rq);
			mat3_mul(Rn, Rq, R);
			mat3_vec(R, m.axis[i], kd.a[i]);
		} else {
#pragma unroll
			for (int k = 0; k < 9; k++) R[k] = Rn[k];
			mat3_vec(R, m.axis[i], kd.a[i]);
			p[0] += kd.a[i][0] * q[i];
			p[1] += kd.a[i][1] * q[i];
			p[2] += kd.a[i][2] * q[i];
		}
#pragma unroll
		for (int k = 0; k < 9; k++) kd.Rb[i][k] = R[k];
		kd.p[i][0] = p[0];
		kd.p[i][1] = p[1];
		kd.p[i][2] = p[2];
	}
}

// Composite rigid body algorithm with spatial inertias expressed about the world origin:
// composite i = bodies i..N-1 as (mass, first moment h, second moment Ibar).
template <int N, bool WITH_GRAVITY>
DEVI void mass_matrix(const DevModel& m, KinDyn<N>& kd) {
	double cm = 0.0, ch[3] = {0, 0, 0};
	double cI[6] = {0, 0, 0, 0, 0, 0};	// xx xy xz yy yz zz
#pragma unroll
	for (int i = N - 1; i >= 0; i--) {
		const double* R = kd.Rb[i];
		// body inertia in world axes: R I R^T
		const double* Ib = m.inertia[i];
		double T[9];  // T = R * I
#pragma unroll
		for (int r = 0; r < 3; r++) {
			T[3 * r + 0] = R[3 * r] * Ib[0] + R[3 * r + 1] * Ib[1] + R[3 * r + 2] * Ib[2];
			T[3 * r + 1] = R[3 * r] * Ib[1] + R[3 * r + 1] * Ib[3] + R[3 * r + 2] * Ib[4];
			T[3 * r + 2] = R[3 * r] * Ib[2] + R[3 * r + 1] * Ib[4] + R[3 * r + 2] * Ib[5];
		}
		double Iw[6];
		Iw[0] = T[0] * R[0] + T[1] * R[1] + T[2] * R[2];
		Iw[1] = T[0] * R[3] + T[1] * R[4] + T[2] * R[5];
		Iw[2] = T[0] * R[6] + T[1] * R[7] + T[2] * R[8];
		Iw[3] = T[3] * R[3] + T[4] * R[4] + T[5] * R[5];
		Iw[4] = T[3] * R[6] + T[4] * R[7] + T[5] * R[8];
		Iw[5] = T[6] * R[6] + T[7] * R[7] + T[8] * R[8];
		double c[3];
		mat3_vec(R, m.com[i], c);
		c[0] += kd.p[i][0];
		c[1] += kd.p[i][1];
		c[2] += kd.p[i][2];
		const double mi = m.mass[i];
		const double cc = dot3(c, c);
		cm += mi;
		ch[0] += mi * c[0];
		ch[1] += mi * c[1];
		ch[2] += mi * c[2];
		cI[0] += Iw[0] + mi * (cc - c[0] * c[0]);
		cI[1] += Iw[1] - mi * c[0] * c[1];
		cI[2] += Iw[2] - mi * c[0] * c[2];
		cI[3] += Iw[3] + mi * (cc - c[1] * c[1]);
		cI[4] += Iw[4] - mi * c[1] * c[2];
		cI[5] += Iw[5] + mi * (cc - c[2] * c[2]);

		// spatial motion of joint i about the world origin: (w, vo)
		double w[3], vo[3];
		if (m.jtype[i] == 0) {
			w[0] = kd.a[i][0];
			w[1] = kd.a[i][1];
			w[2] = kd.a[i][2];
			cross3(kd.p[i], kd.a[i], vo);
		} else {
			w[0] = w[1] = w[2] = 0.0;
			vo[0] = kd.a[i][0];
			vo[1] = kd.a[i][1];
			vo[2] = kd.a[i][2];
		}
		// spatial momentum of the composite under that motion: linear f, angular no (about origin)
		double f[3], no[3], t1[3];
		cross3(w, ch, t1);
		f[0] = cm * vo[0] + t1[0];
		f[1] = cm * vo[1] + t1[1];
		f[2] = cm * vo[2] + t1[2];
		cross3(ch, vo, t1);
		no[0] = cI[0] * w[0] + cI[1] * w[1] + cI[2] * w[2] + t1[0];
		no[1] = cI[1] * w[0] + cI[3] * w[1] + cI[4] * w[2] + t1[1];
		no[2] = cI[2] * w[0] + cI[4] * w[1] + cI[5] * w[2] + t1[2];
#pragma unroll
		for (int j = 0; j <= i; j++) {
			double wj[3], voj[3];
			if (m.jtype[j] == 0) {
				wj[0] = kd.a[j][0];
				wj[1] = kd.a[j][1];
				wj[2] = kd.a[j][2];
				cross3(kd.p[j], kd.a[j], voj);
			} else {
				wj[0] = wj[1] = wj[2] = 0.0;
				voj[0] = kd.a[j][0];
				voj[1] = kd.a[j][1];
				voj[2] = kd.a[j][2];
			}
			const double v = dot3(wj, no) + dot3(voj, f);
			kd.M[i][j] = v;
			kd.M[j][i] = v;
		}
		if (WITH_GRAVITY) {
			// -(s_i . gravity wrench of the composite about the origin)
			double hg[3];
			cross3(ch, m.gravity, hg);
			kd.g[i] = -(dot3(w, hg) + cm * dot3(vo, m.gravity));
		}
	}
}

// Bias forces b(q, dq) + g(q) by the recursive Newton-Euler algorithm in world coordinates (the
// simulation side of the loop; restated in oracle/simulation.py).  The base is at rest and "accelerates" with -gravity.
template <int N>
DEVI void rnea_bias(const DevModel& m, const KinDyn<N>& kd, const double (&dq)[N], double (&tau)[N]) {
	double W[N][3], AL[N][3], AP[N][3];
	double w[3] = {0, 0, 0}, al[3] = {0, 0, 0}, ap[3] = {-m.gravity[0], -m.gravity[1], -m.gravity[2]};
	double pp[3] = {0, 0, 0};
#pragma unroll
	for (int k = 0; k < N; k++) {
		const double r[3] = {kd.p[k][0] - pp[0], kd.p[k][1] - pp[1], kd.p[k][2] - pp[2]};
		double t1[3], t2[3];
		cross3(al, r, t1);
		cross3(w, r, t2);
		double t3[3];
		cross3(w, t2, t3);
#pragma unroll
		for (int c = 0; c < 3; c++) ap[c] += t1[c] + t3[c];
		cross3(w, kd.a[k], t1);
		if (m.jtype[k] == 0) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				al[c] += t1[c] * dq[k];
				w[c] += kd.a[k][c] * dq[k];
			}
		} else {
#pragma unroll
			for (int c = 0; c < 3; c++) ap[c] += 2.0 * t1[c] * dq[k];
		}
#pragma unroll
		for (int c = 0; c < 3; c++) {
			W[k][c] = w[c];
			AL[k][c] = al[c];
			AP[k][c] = ap[c];
			pp[c] = kd.p[k][c];
		}
	}
	double f[3] = {0, 0, 0}, nm[3] = {0, 0, 0};
#pragma unroll
	for (int k = N - 1; k >= 0; k--) {
		const double* R = kd.Rb[k];
		double cw[3];
		mat3_vec(R, m.com[k], cw);
		// I_w x = R (I (R^T x))
		auto apply_inertia = [&](const double x[3], double y[3]) {
			double l[3], il[3];
			mat3t_vec(R, x, l);
			const double* Ib = m.inertia[k];
			il[0] = Ib[0] * l[0] + Ib[1] * l[1] + Ib[2] * l[2];
			il[1] = Ib[1] * l[0] + Ib[3] * l[1] + Ib[4] * l[2];
			il[2] = Ib[2] * l[0] + Ib[4] * l[1] + Ib[5] * l[2];
			mat3_vec(R, il, y);
		};
		double t1[3], t2[3], ac[3], F[3], Nc[3], Iw_w[3], Iw_al[3];
		cross3(AL[k], cw, t1);
		cross3(W[k], cw, t2);
		double t3[3];
		cross3(W[k], t2, t3);
#pragma unroll
		for (int c = 0; c < 3; c++) {
			ac[c] = AP[k][c] + t1[c] + t3[c];
			F[c] = m.mass[k] * ac[c];
		}
		apply_inertia(W[k], Iw_w);
		apply_inertia(AL[k], Iw_al);
		cross3(W[k], Iw_w, t1);
#pragma unroll
		for (int c = 0; c < 3; c++) Nc[c] = Iw_al[c] + t1[c];
		if (k < N - 1) {
			const double d[3] = {kd.p[k + 1 < N ? k + 1 : k][0] - kd.p[k][0], kd.p[k + 1 < N ? k + 1 : k][1] - kd.p[k][1],
								 kd.p[k + 1 < N ? k + 1 : k][2] - kd.p[k][2]};
			cross3(d, f, t2);
#pragma unroll
			for (int c = 0; c < 3; c++) nm[c] += t2[c];
		}
		cross3(cw, F, t2);
#pragma unroll
		for (int c = 0; c < 3; c++) {
			nm[c] += Nc[c] + t2[c];
			f[c] += F[c];
		}
		tau[k] = (m.jtype[k] == 0) ? dot3(kd.a[k], nm) : dot3(kd.a[k], f);
	}
}

// 6 x n world-frame Jacobian (linear rows first) of point x fixed to body `body`,
// stored transposed: JT[i][0..2] = linear column of joint i, JT[i][3..5] = angular.
template <int N>
DEVI void point_jacobian_t(const DevModel& m, const KinDyn<N>& kd, int body, const double x[3], double (&JT)[N][6]) {
#pragma unroll
	for (int i = 0; i < N; i++) {
		if (i <= body) {
			if (m.jtype[i] == 0) {
				double d[3] = {x[0] - kd.p[i][0], x[1] - kd.p[i][1], x[2] - kd.p[i][2]};
				double v[3];
				cross3(kd.a[i], d, v);
				JT[i][0] = v[0];
				JT[i][1] = v[1];
				JT[i][2] = v[2];
				JT[i][3] = kd.a[i][0];
				JT[i][4] = kd.a[i][1];
				JT[i][5] = kd.a[i][2];
			} else {
				JT[i][0] = kd.a[i][0];
				JT[i][1] = kd.a[i][1];
				JT[i][2] = kd.a[i][2];
				JT[i][3] = JT[i][4] = JT[i][5] = 0.0;
			}
		} else {
#pragma unroll
			for (int k = 0; k < 6; k++) JT[i][k] = 0.0;
		}
	}
}

// pose of a frame (Rf, tf given in the body frame) in the world
template <int N>
DEVI void frame_pose(const KinDyn<N>& kd, int body, const double Rf[9], const double tf[3], double x[3], double R[9]) {
	// body is warp-uniform but not a compile-time constant: select without dynamic indexing
	double Rb[9], pb[3];
#pragma unroll
	for (int k = 0; k < 9; k++) Rb[k] = (k % 4 == 0) ? 1.0 : 0.0;	// body < 0: frame fixed to the world
	pb[0] = pb[1] = pb[2] = 0.0;
#pragma unroll
	for (int i = 0; i < N; i++) {
		if (i == body) {
#pragma unroll
			for (int k = 0; k < 9; k++) Rb[k] = kd.Rb[i][k];
			pb[0] = kd.p[i][0];
			pb[1] = kd.p[i][1];
			pb[2] = kd.p[i][2];
		}
	}
	double t[3];
	mat3_vec(Rb, tf, t);
	x[0] = pb[0] + t[0];
	x[1] = pb[1] + t[1];
	x[2] = pb[2] + t[2];
	mat3_mul(Rb, Rf, R);
}

// ------------------------------------------------------------------------------------------------
// Variant used by the fused cycle kernel: the body orientations produced by the forward pass are staged in
// shared memory (element e of this thread at smt[e * sms]) instead of 9 N registers, joints whose axis is the
// local z axis and bodies with isotropic inertia take short cuts (warp-uniform branches on model constants),
// and only the lower triangle of M is written.
template <int N>
struct KinDynS {
	double a[N][3];	 // joint axis, world
	double p[N][3];	 // joint origin, world
	double M[N][N];	 // lower triangle valid
	double g[N];
};

// ZREV: every joint is revolute about its local z axis (checked on the host: model_all_revolute_z), so the per-joint
// alternatives disappear from the instruction stream.
template <int N, bool ZREV = false>
DEVI void forward_kinematics_s(const DevModel& m, const double (&q)[N], KinDynS<N>& kd, double* smt, int sms) {
	double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	double p[3] = {0, 0, 0};
	// all joint angles first: N independent polynomial chains instead of one per link of the serial chain below
	double sn[N], cs[N];
#pragma unroll
	for (int i = 0; i < N; i++) {
		sn[i] = 0.0;
		cs[i] = 1.0;
		if (ZREV || m.jtype[i] == 0) sincos_joint(q[i], &sn[i], &cs[i]);
	}
#pragma unroll
	for (int i = 0; i < N; i++) {
		double t[3];
		mat3_vec(R, m.t_fix[i], t);
		p[0] += t[0];
		p[1] += t[1];
		p[2] += t[2];
		double Rn[9];
		mat3_mul(R, m.R_fix[i], Rn);
		const bool axis_z = ZREV || ((m.axis[i][0] == 0.0) && (m.axis[i][1] == 0.0) && (m.axis[i][2] == 1.0));
		if (ZREV || m.jtype[i] == 0) {
			const double s = sn[i], c = cs[i];
			if (axis_z) {  // R = Rn Rz(q): only the first two columns mix
#pragma unroll
				for (int r = 0; r < 3; r++) {
					const double c0 = Rn[3 * r], c1 = Rn[3 * r + 1];
					R[3 * r] = c * c0 + s * c1;
					R[3 * r + 1] = c * c1 - s * c0;
					R[3 * r + 2] = Rn[3 * r + 2];
				}
				kd.a[i][0] = R[2];
				kd.a[i][1] = R[5];
				kd.a[i][2] = R[8];
			} else {
				double Rq[9];
				axis_angle(m.axis[i], s, c, Rq);
				mat3_mul(Rn, Rq, R);
				mat3_vec(R, m.axis[i], kd.a[i]);
			}
		} else {
#pragma unroll
			for (int k = 0; k < 9; k++) R[k] = Rn[k];
			mat3_vec(R, m.axis[i], kd.a[i]);
			p[0] += kd.a[i][0] * q[i];
			p[1] += kd.a[i][1] * q[i];
			p[2] += kd.a[i][2] * q[i];
		}
#pragma unroll
		for (int k = 0; k < 9; k++) smt[(9 * i + k) * sms] = R[k];
		kd.p[i][0] = p[0];
		kd.p[i][1] = p[1];
		kd.p[i][2] = p[2];
		OSC_LS();
	}
}

template <int N, bool WITH_GRAVITY, bool ZREV = false>
DEVI void mass_matrix_s(const DevModel& m, KinDynS<N>& kd, const double* smt, int sms) {
	double cm = 0.0, ch[3] = {0, 0, 0};
	double cI[6] = {0, 0, 0, 0, 0, 0};	// xx xy xz yy yz zz
#pragma unroll
	for (int i = N - 1; i >= 0; i--) {
		double R[9];
#pragma unroll
		for (int k = 0; k < 9; k++) R[k] = smt[(9 * i + k) * sms];
		const double* Ib = m.inertia[i];
		const bool iso = (Ib[1] == 0.0) && (Ib[2] == 0.0) && (Ib[4] == 0.0) && (Ib[0] == Ib[3]) && (Ib[0] == Ib[5]);
		double Iw[6];
		if (iso) {	// R (k I) R^T = k I
			Iw[0] = Ib[0];
			Iw[1] = 0.0;
			Iw[2] = 0.0;
			Iw[3] = Ib[0];
			Iw[4] = 0.0;
			Iw[5] = Ib[0];
		} else {
			double T[9];  // T = R * I
#pragma unroll
			for (int r = 0; r < 3; r++) {
				T[3 * r + 0] = R[3 * r] * Ib[0] + R[3 * r + 1] * Ib[1] + R[3 * r + 2] * Ib[2];
				T[3 * r + 1] = R[3 * r] * Ib[1] + R[3 * r + 1] * Ib[3] + R[3 * r + 2] * Ib[4];
				T[3 * r + 2] = R[3 * r] * Ib[2] + R[3 * r + 1] * Ib[4] + R[3 * r + 2] * Ib[5];
			}
			Iw[0] = T[0] * R[0] + T[1] * R[1] + T[2] * R[2];
			Iw[1] = T[0] * R[3] + T[1] * R[4] + T[2] * R[5];
			Iw[2] = T[0] * R[6] + T[1] * R[7] + T[2] * R[8];
			Iw[3] = T[3] * R[3] + T[4] * R[4] + T[5] * R[5];
			Iw[4] = T[3] * R[6] + T[4] * R[7] + T[5] * R[8];
			Iw[5] = T[6] * R[6] + T[7] * R[7] + T[8] * R[8];
		}
		double c[3];
		mat3_vec(R, m.com[i], c);
		c[0] += kd.p[i][0];
		c[1] += kd.p[i][1];
		c[2] += kd.p[i][2];
		const double mi = m.mass[i];
		const double cc = dot3(c, c);
		cm += mi;
		ch[0] += mi * c[0];
		ch[1] += mi * c[1];
		ch[2] += mi * c[2];
		cI[0] += Iw[0] + mi * (cc - c[0] * c[0]);
		cI[1] += Iw[1] - mi * c[0] * c[1];
		cI[2] += Iw[2] - mi * c[0] * c[2];
		cI[3] += Iw[3] + mi * (cc - c[1] * c[1]);
		cI[4] += Iw[4] - mi * c[1] * c[2];
		cI[5] += Iw[5] + mi * (cc - c[2] * c[2]);

		double w[3], vo[3];
		if (ZREV || m.jtype[i] == 0) {
			w[0] = kd.a[i][0];
			w[1] = kd.a[i][1];
			w[2] = kd.a[i][2];
			cross3(kd.p[i], kd.a[i], vo);
		} else {
			w[0] = w[1] = w[2] = 0.0;
			vo[0] = kd.a[i][0];
			vo[1] = kd.a[i][1];
			vo[2] = kd.a[i][2];
		}
		double f[3], no[3], t1[3];
		cross3(w, ch, t1);
		f[0] = cm * vo[0] + t1[0];
		f[1] = cm * vo[1] + t1[1];
		f[2] = cm * vo[2] + t1[2];
		cross3(ch, vo, t1);
		no[0] = cI[0] * w[0] + cI[1] * w[1] + cI[2] * w[2] + t1[0];
		no[1] = cI[1] * w[0] + cI[3] * w[1] + cI[4] * w[2] + t1[1];
		no[2] = cI[2] * w[0] + cI[4] * w[1] + cI[5] * w[2] + t1[2];
#pragma unroll
		for (int j = 0; j <= i; j++) {
			// s_j . F_i = a_j . (n_o + f x p_j) for a revolute joint j,  a_j . f for a prismatic one
			if (ZREV || m.jtype[j] == 0) {
				double fxp[3];
				cross3(f, kd.p[j], fxp);
				kd.M[i][j] = kd.a[j][0] * (no[0] + fxp[0]) + kd.a[j][1] * (no[1] + fxp[1]) + kd.a[j][2] * (no[2] + fxp[2]);
			} else {
				kd.M[i][j] = dot3(kd.a[j], f);
			}
		}
		if (WITH_GRAVITY) {
			double hg[3];
			cross3(ch, m.gravity, hg);
			kd.g[i] = -(dot3(w, hg) + cm * dot3(vo, m.gravity));
		}
		OSC_LS();
	}
}

// ------------------------------------------------------------------------------------------------
// Rolled kinematics / dynamics for chains whose joints are all revolute about their local z axis (the SPEC
// instantiation of the fused kernel).  The loops over the joints are real loops: the per-joint quantities live in
// shared memory under a run-time index and the model constants come from the constant bank under a uniform index,
// so this stage costs about 500 instructions of code instead of 2,400 (the instruction stream of the whole kernel
// has to stay inside the instruction cache, profiles/r01_ifetch.md).  Shared-memory slots of a thread, in doubles:
//   [0, 9 N)      body orientations R_j, row major; row i of the mass matrix later takes the place of R_i
//   [9 N, 15 N)   joint axis a_j (3) and joint origin p_j (3), world
template <int N>
constexpr int kSmAxes = 9 * N;

// forward kinematics; the joint angles are parked in the first slot of the orientation they produce (slot 9 i), so
// that the loop reads them under its run-time index
template <int N>
DEVI void forward_kinematics_rolled(const DevModel& m, const double (&qr)[N], double* smt, int sms) {
	double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	double p[3] = {0, 0, 0};
#pragma unroll
	for (int i = 0; i < N; i++) smt[(size_t)(9 * i) * sms] = qr[i];
	// sine and cosine of joint i + 1 are evaluated while the orientation of body i goes through its dependent chain
	double sn, cn;
	sincos_joint(qr[0], &sn, &cn);
#pragma unroll 1
	for (int i = 0; i < N; i++) {
		const double s = sn, c = cn;
		if (i + 1 < N) sincos_joint(smt[(size_t)(9 * (i + 1)) * sms], &sn, &cn);
		double t[3];
		mat3_vec(R, m.t_fix[i], t);
		p[0] += t[0];
		p[1] += t[1];
		p[2] += t[2];
		double Rn[9];
		mat3_mul(R, m.R_fix[i], Rn);
#pragma unroll
		for (int r = 0; r < 3; r++) {  // R = Rn Rz(q): only the first two columns mix
			const double c0 = Rn[3 * r], c1 = Rn[3 * r + 1];
			R[3 * r] = c * c0 + s * c1;
			R[3 * r + 1] = c * c1 - s * c0;
			R[3 * r + 2] = Rn[3 * r + 2];
		}
		double* Rs = smt + (size_t)(9 * i) * sms;
#pragma unroll
		for (int k = 0; k < 9; k++) Rs[k * sms] = R[k];
		double* ap = smt + (size_t)(kSmAxes<N> + 6 * i) * sms;
		ap[0] = R[2];
		ap[1 * sms] = R[5];
		ap[2 * sms] = R[8];
		ap[3 * sms] = p[0];
		ap[4 * sms] = p[1];
		ap[5 * sms] = p[2];
	}
}

// composite-rigid-body mass matrix; on exit M(i, j), j <= i, is at slot 9 i + j and, WITH_GRAVITY, entry i of the
// gravity vector at slot 9 i + 8
template <int N, bool WITH_GRAVITY = false>
DEVI void mass_matrix_rolled(const DevModel& m, double* smt, int sms) {
	double a[N][3], p[N][3];
#pragma unroll
	for (int j = 0; j < N; j++)
#pragma unroll
		for (int k = 0; k < 3; k++) {
			a[j][k] = smt[(size_t)(kSmAxes<N> + 6 * j + k) * sms];
			p[j][k] = smt[(size_t)(kSmAxes<N> + 6 * j + 3 + k) * sms];
		}
	double cm = 0.0, ch[3] = {0, 0, 0};
	double cI[6] = {0, 0, 0, 0, 0, 0};	// xx xy xz yy yz zz
#pragma unroll 1
	for (int i = N - 1; i >= 0; i--) {
		double* Rs = smt + (size_t)(9 * i) * sms;
		double R[9];
#pragma unroll
		for (int k = 0; k < 9; k++) R[k] = Rs[k * sms];
		const double* ap = smt + (size_t)(kSmAxes<N> + 6 * i) * sms;
		const double ai[3] = {ap[0], ap[1 * sms], ap[2 * sms]};
		const double pi[3] = {ap[3 * sms], ap[4 * sms], ap[5 * sms]};
		const double* Ib = m.inertia[i];
		const bool iso = (Ib[1] == 0.0) && (Ib[2] == 0.0) && (Ib[4] == 0.0) && (Ib[0] == Ib[3]) && (Ib[0] == Ib[5]);
		double Iw[6];
		if (iso) {	// R (k I) R^T = k I  (warp-uniform: a model constant)
			Iw[0] = Ib[0];
			Iw[1] = 0.0;
			Iw[2] = 0.0;
			Iw[3] = Ib[0];
			Iw[4] = 0.0;
			Iw[5] = Ib[0];
		} else {
			double T[9];  // T = R * I
#pragma unroll
			for (int r = 0; r < 3; r++) {
				T[3 * r + 0] = R[3 * r] * Ib[0] + R[3 * r + 1] * Ib[1] + R[3 * r + 2] * Ib[2];
				T[3 * r + 1] = R[3 * r] * Ib[1] + R[3 * r + 1] * Ib[3] + R[3 * r + 2] * Ib[4];
				T[3 * r + 2] = R[3 * r] * Ib[2] + R[3 * r + 1] * Ib[4] + R[3 * r + 2] * Ib[5];
			}
			Iw[0] = T[0] * R[0] + T[1] * R[1] + T[2] * R[2];
			Iw[1] = T[0] * R[3] + T[1] * R[4] + T[2] * R[5];
			Iw[2] = T[0] * R[6] + T[1] * R[7] + T[2] * R[8];
			Iw[3] = T[3] * R[3] + T[4] * R[4] + T[5] * R[5];
			Iw[4] = T[3] * R[6] + T[4] * R[7] + T[5] * R[8];
			Iw[5] = T[6] * R[6] + T[7] * R[7] + T[8] * R[8];
		}
		double c[3];
		mat3_vec(R, m.com[i], c);
		c[0] += pi[0];
		c[1] += pi[1];
		c[2] += pi[2];
		const double mi = m.mass[i];
		const double cc = dot3(c, c);
		cm += mi;
		ch[0] += mi * c[0];
		ch[1] += mi * c[1];
		ch[2] += mi * c[2];
		cI[0] += Iw[0] + mi * (cc - c[0] * c[0]);
		cI[1] += Iw[1] - mi * c[0] * c[1];
		cI[2] += Iw[2] - mi * c[0] * c[2];
		cI[3] += Iw[3] + mi * (cc - c[1] * c[1]);
		cI[4] += Iw[4] - mi * c[1] * c[2];
		cI[5] += Iw[5] + mi * (cc - c[2] * c[2]);
		// spatial momentum of the composite under a unit velocity of joint i (w = a_i, v_o = p_i x a_i)
		double vo[3], f[3], no[3], t1[3];
		cross3(pi, ai, vo);
		cross3(ai, ch, t1);
		f[0] = cm * vo[0] + t1[0];
		f[1] = cm * vo[1] + t1[1];
		f[2] = cm * vo[2] + t1[2];
		cross3(ch, vo, t1);
		no[0] = cI[0] * ai[0] + cI[1] * ai[1] + cI[2] * ai[2] + t1[0];
		no[1] = cI[1] * ai[0] + cI[3] * ai[1] + cI[4] * ai[2] + t1[1];
		no[2] = cI[2] * ai[0] + cI[4] * ai[1] + cI[5] * ai[2] + t1[2];
		if constexpr (WITH_GRAVITY) {	// -(s_i . gravity wrench of the composite about the origin)
			double hg[3];
			cross3(ch, m.gravity, hg);
			Rs[8 * sms] = -(dot3(ai, hg) + cm * dot3(vo, m.gravity));
		}
		// M(i, j) = a_j . (n_o + f x p_j) for j <= i: the loop over j is unrolled (compile-time register indices) and the
		// entries above the diagonal are skipped by a warp-uniform test
#pragma unroll
		for (int j = 0; j < N; j++) {
			if (j <= i) {
				double fxp[3];
				cross3(f, p[j], fxp);
				Rs[j * sms] = a[j][0] * (no[0] + fxp[0]) + a[j][1] * (no[1] + fxp[1]) + a[j][2] * (no[2] + fxp[2]);
			}
		}
	}
}

// pose of a frame (Rf, tf given in the body frame) in the world from the rolled layout; body >= 0
template <int N>
DEVI void frame_pose_rolled(int body, const double Rf[9], const double tf[3], double x[3], double R[9], const double* smt, int sms) {
	double Rb[9], pb[3];
	const double* Rs = smt + (size_t)(9 * body) * sms;
#pragma unroll
	for (int k = 0; k < 9; k++) Rb[k] = Rs[k * sms];
	const double* ap = smt + (size_t)(kSmAxes<N> + 6 * body) * sms;
	pb[0] = ap[3 * sms];
	pb[1] = ap[4 * sms];
	pb[2] = ap[5 * sms];
	double t[3];
	mat3_vec(Rb, tf, t);
	x[0] = pb[0] + t[0];
	x[1] = pb[1] + t[1];
	x[2] = pb[2] + t[2];
	mat3_mul(Rb, Rf, R);
}

// pose of a frame (Rf, tf given in the body frame) in the world; body orientation read back from shared memory
template <int N>
DEVI void frame_pose_s(const KinDynS<N>& kd, int body, const double Rf[9], const double tf[3], double x[3], double R[9],
						const double* smt, int sms) {
	double Rb[9], pb[3];
	if (body >= 0) {
#pragma unroll
		for (int k = 0; k < 9; k++) Rb[k] = smt[(9 * body + k) * sms];
	} else {
#pragma unroll
		for (int k = 0; k < 9; k++) Rb[k] = (k % 4 == 0) ? 1.0 : 0.0;
	}
	pb[0] = pb[1] = pb[2] = 0.0;
#pragma unroll
	for (int i = 0; i < N; i++) {
		if (i == body) {
			pb[0] = kd.p[i][0];
			pb[1] = kd.p[i][1];
			pb[2] = kd.p[i][2];
		}
	}
	double t[3];
	mat3_vec(Rb, tf, t);
	x[0] = pb[0] + t[0];
	x[1] = pb[1] + t[1];
	x[2] = pb[2] + t[2];
	mat3_mul(Rb, Rf, R);
}

// column j of the 6 x n world-frame Jacobian (linear rows first) of point x fixed to body `body`
template <int N, bool ZREV = false>
DEVI void jacobian_column(const DevModel& m, const KinDynS<N>& kd, int body, const double x[3], int j, double c6[6]) {
	if (j <= body) {
		if (ZREV || m.jtype[j] == 0) {
			double d[3] = {x[0] - kd.p[j][0], x[1] - kd.p[j][1], x[2] - kd.p[j][2]};
			double v[3];
			cross3(kd.a[j], d, v);
			c6[0] = v[0];
			c6[1] = v[1];
			c6[2] = v[2];
			c6[3] = kd.a[j][0];
			c6[4] = kd.a[j][1];
			c6[5] = kd.a[j][2];
		} else {
			c6[0] = kd.a[j][0];
			c6[1] = kd.a[j][1];
			c6[2] = kd.a[j][2];
			c6[3] = c6[4] = c6[5] = 0.0;
		}
	} else {
#pragma unroll
		for (int k = 0; k < 6; k++) c6[k] = 0.0;
	}
}

// Pose of a frame fixed to body `body` for joint positions q, without keeping any per-joint data (used by the
// singularity classification, which only needs the perturbed end pose: SingularityHandler.cpp:255-260).
template <int N>
DEVI void pose_only(const DevModel& m, const double (&q)[N], int body, const double Rf[9], const double tf[3], double x[3], double Rout[9]) {
	double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	double p[3] = {0, 0, 0};
#pragma unroll
	for (int i = 0; i < N; i++) {
		if (i <= body) {
			double t[3];
			mat3_vec(R, m.t_fix[i], t);
			p[0] += t[0];
			p[1] += t[1];
			p[2] += t[2];
			double Rn[9];
			mat3_mul(R, m.R_fix[i], Rn);
			if (m.jtype[i] == 0) {
				double s, c;
				sincos(q[i], &s, &c);
				double Rq[9];
				axis_angle(m.axis[i], s, c, Rq);
				mat3_mul(Rn, Rq, R);
			} else {
				double a[3];
#pragma unroll
				for (int k = 0; k < 9; k++) R[k] = Rn[k];
				mat3_vec(R, m.axis[i], a);
				p[0] += a[0] * q[i];
				p[1] += a[1] * q[i];
				p[2] += a[2] * q[i];
			}
		}
	}
	double t[3];
	mat3_vec(R, tf, t);
	x[0] = p[0] + t[0];
	x[1] = p[1] + t[1];
	x[2] = p[2] + t[2];
	mat3_mul(R, Rf, Rout);
}

}  // namespace osc
