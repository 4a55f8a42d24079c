// ORACLE / CPU BASELINE (test infrastructure only) -- plain C++17 restatement of the reference
// hot path, no Eigen, no sai-model (neither exists in this environment).
//
// PARITY UNPINNED: the reference cannot be compiled here (it needs Eigen3 and sai-model@master,
// CMakeLists.txt:28-33) and ships no tests; this file restates its source statement by statement,
// keeping the reference's per-cycle structure on purpose -- dynamic heap matrices, a thin SVD of the
// projected Jacobian in every SingularityHandler::updateTaskModel, M_BIE.inverse() recomputed in every
// task, the Jacobian evaluated in updateTaskModel and again in computeTorques -- because it is also the
// CPU baseline bench.py times beside the CUDA path.  It is validated against the
// independent numpy restatement (oracle/primitives.py) by tests/test_cpp_oracle.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py may load the library built from this file.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <stdexcept>
#include <thread>
#include <vector>

namespace oref {

// ------------------------------------------------------------------ dense matrices (MatrixXd stand-in)
struct Mat {
	int r = 0, c = 0;
	std::vector<double> a;
	Mat() {}
	Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
	double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
	double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
	static Mat identity(int n) {
		Mat m(n, n);
		for (int i = 0; i < n; i++) m(i, i) = 1.0;
		return m;
	}
	double norm() const {
		double s = 0;
		for (double v : a) s += v * v;
		return std::sqrt(s);
	}
};
typedef std::vector<double> Vec;

Mat operator*(const Mat& A, const Mat& B) {
	Mat C(A.r, B.c);
	for (int i = 0; i < A.r; i++)
		for (int k = 0; k < A.c; k++) {
			const double v = A(i, k);
			if (v == 0.0) continue;
			for (int j = 0; j < B.c; j++) C(i, j) += v * B(k, j);
		}
	return C;
}
Vec operator*(const Mat& A, const Vec& x) {
	Vec y(A.r, 0.0);
	for (int i = 0; i < A.r; i++) {
		double s = 0;
		for (int j = 0; j < A.c; j++) s += A(i, j) * x[j];
		y[i] = s;
	}
	return y;
}
Mat operator-(const Mat& A, const Mat& B) {
	Mat C(A.r, A.c);
	for (size_t i = 0; i < A.a.size(); i++) C.a[i] = A.a[i] - B.a[i];
	return C;
}
Mat T(const Mat& A) {
	Mat B(A.c, A.r);
	for (int i = 0; i < A.r; i++)
		for (int j = 0; j < A.c; j++) B(j, i) = A(i, j);
	return B;
}
Vec operator+(const Vec& a, const Vec& b) {
	Vec c(a.size());
	for (size_t i = 0; i < a.size(); i++) c[i] = a[i] + b[i];
	return c;
}
Vec operator-(const Vec& a, const Vec& b) {
	Vec c(a.size());
	for (size_t i = 0; i < a.size(); i++) c[i] = a[i] - b[i];
	return c;
}
Vec operator*(double s, const Vec& a) {
	Vec c(a.size());
	for (size_t i = 0; i < a.size(); i++) c[i] = s * a[i];
	return c;
}
double dot(const Vec& a, const Vec& b) {
	double s = 0;
	for (size_t i = 0; i < a.size(); i++) s += a[i] * b[i];
	return s;
}
double norm(const Vec& a) { return std::sqrt(dot(a, a)); }
Mat block_cols(const Mat& A, int c0, int nc) {
	Mat B(A.r, nc);
	for (int i = 0; i < A.r; i++)
		for (int j = 0; j < nc; j++) B(i, j) = A(i, c0 + j);
	return B;
}
Mat block_rows(const Mat& A, int r0, int nr) {
	Mat B(nr, A.c);
	for (int i = 0; i < nr; i++)
		for (int j = 0; j < A.c; j++) B(i, j) = A(r0 + i, j);
	return B;
}
Vec col(const Mat& A, int j) {
	Vec v(A.r);
	for (int i = 0; i < A.r; i++) v[i] = A(i, j);
	return v;
}

// Eigen dynamic .inverse(): LU with partial pivoting
Mat inverse(const Mat& A_in) {
	const int n = A_in.r;
	Mat A = A_in, B = Mat::identity(n);
	for (int k = 0; k < n; k++) {
		int p = k;
		for (int i = k + 1; i < n; i++)
			if (std::fabs(A(i, k)) > std::fabs(A(p, k))) p = i;
		if (p != k)
			for (int j = 0; j < n; j++) {
				std::swap(A(k, j), A(p, j));
				std::swap(B(k, j), B(p, j));
			}
		const double d = 1.0 / A(k, k);
		for (int j = 0; j < n; j++) {
			A(k, j) *= d;
			B(k, j) *= d;
		}
		for (int i = 0; i < n; i++) {
			if (i == k) continue;
			const double f = A(i, k);
			if (f == 0.0) continue;
			for (int j = 0; j < n; j++) {
				A(i, j) -= f * A(k, j);
				B(i, j) -= f * B(k, j);
			}
		}
	}
	return B;
}

// A.llt().solve(Identity)
Mat llt_inverse(const Mat& A) {
	const int n = A.r;
	Mat L(n, n);
	for (int j = 0; j < n; j++) {
		double d = A(j, j);
		for (int k = 0; k < j; k++) d -= L(j, k) * L(j, k);
		L(j, j) = std::sqrt(d);
		for (int i = j + 1; i < n; i++) {
			double s = A(i, j);
			for (int k = 0; k < j; k++) s -= L(i, k) * L(j, k);
			L(i, j) = s / L(j, j);
		}
	}
	Mat X = Mat::identity(n);
	for (int c = 0; c < n; c++) {
		for (int i = 0; i < n; i++) {
			double s = X(i, c);
			for (int k = 0; k < i; k++) s -= L(i, k) * X(k, c);
			X(i, c) = s / L(i, i);
		}
		for (int i = n - 1; i >= 0; i--) {
			double s = X(i, c);
			for (int k = i + 1; k < n; k++) s -= L(k, i) * X(k, c);
			X(i, c) = s / L(i, i);
		}
	}
	return X;
}

// Thin SVD by one-sided Jacobi (Hestenes), singular values sorted descending (JacobiSVD stand-in).
struct SVD {
	Mat U, V;
	Vec s;
};
SVD svd_thin(const Mat& A_in) {
	const bool transposed = A_in.r < A_in.c;
	Mat A = transposed ? T(A_in) : A_in;  // m >= n
	const int m = A.r, n = A.c;
	Mat V = Mat::identity(n);
	for (int sweep = 0; sweep < 60; sweep++) {
		bool rotated = false;
		for (int p = 0; p < n - 1; p++)
			for (int q = p + 1; q < n; q++) {
				double alpha = 0, beta = 0, gamma = 0;
				for (int i = 0; i < m; i++) {
					alpha += A(i, p) * A(i, p);
					beta += A(i, q) * A(i, q);
					gamma += A(i, p) * A(i, q);
				}
				if (std::fabs(gamma) <= 1e-16 * std::sqrt(alpha * beta) || gamma == 0.0) continue;
				rotated = true;
				const double zeta = (beta - alpha) / (2.0 * gamma);
				const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
				const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
				for (int i = 0; i < m; i++) {
					const double ap = A(i, p), aq = A(i, q);
					A(i, p) = c * ap - s * aq;
					A(i, q) = s * ap + c * aq;
				}
				for (int i = 0; i < n; i++) {
					const double vp = V(i, p), vq = V(i, q);
					V(i, p) = c * vp - s * vq;
					V(i, q) = s * vp + c * vq;
				}
			}
		if (!rotated) break;
	}
	Vec s(n);
	for (int j = 0; j < n; j++) {
		double nn = 0;
		for (int i = 0; i < m; i++) nn += A(i, j) * A(i, j);
		s[j] = std::sqrt(nn);
	}
	std::vector<int> order(n);
	for (int j = 0; j < n; j++) order[j] = j;
	std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return s[a] > s[b]; });
	Mat Us(m, n), Vs(n, n);
	Vec ss(n);
	for (int jj = 0; jj < n; jj++) {
		const int j = order[jj];
		ss[jj] = s[j];
		const double inv = s[j] > 0 ? 1.0 / s[j] : 0.0;
		for (int i = 0; i < m; i++) Us(i, jj) = A(i, j) * inv;
		for (int i = 0; i < n; i++) Vs(i, jj) = V(i, j);
	}
	SVD out;
	out.s = ss;
	if (transposed) {
		out.U = Vs;
		out.V = Us;
	} else {
		out.U = Us;
		out.V = Vs;
	}
	// sign convention of this repo: largest-magnitude entry of every V column positive
	for (int j = 0; j < out.V.c; j++) {
		int im = 0;
		for (int i = 1; i < out.V.r; i++)
			if (std::fabs(out.V(i, j)) > std::fabs(out.V(im, j))) im = i;
		if (out.V(im, j) < 0) {
			for (int i = 0; i < out.V.r; i++) out.V(i, j) = -out.V(i, j);
			for (int i = 0; i < out.U.r; i++) out.U(i, j) = -out.U(i, j);
		}
	}
	return out;
}

Mat pinv(const Mat& A, double rel_tol) {
	SVD d = svd_thin(A);
	const int k = (int)d.s.size();
	Mat S(k, k);
	for (int i = 0; i < k; i++) S(i, i) = (d.s[i] > rel_tol * d.s[0] && d.s[i] > 0) ? 1.0 / d.s[i] : 0.0;
	return d.V * S * T(d.U);
}

// SaiModel::matrixRangeBasis
Mat matrixRangeBasis(const Mat& A, double tol = 1e-3) {
	const int rows = A.r;
	if (A.norm() < tol) return Mat(rows, 1);
	SVD d = svd_thin(A);
	if (d.s[0] < tol) return Mat(rows, 1);
	int task_dof = std::min(A.r, A.c);
	for (int i = (int)d.s.size() - 1; i > 0; i--) {
		if (d.s[i] / d.s[0] < tol)
			task_dof--;
		else
			break;
	}
	if (task_dof == rows) return Mat::identity(rows);
	return block_cols(d.U, 0, task_dof);
}

// ------------------------------------------------------------------ 3-vectors / rotations
struct V3 {
	double x = 0, y = 0, z = 0;
};
struct M3 {
	double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
};
V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
V3 operator*(const M3& A, V3 v) {
	return {A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
			A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
M3 operator*(const M3& A, const M3& B) {
	M3 C;
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
	return C;
}
M3 Tr(const M3& A) {
	M3 B;
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) B.m[3 * i + j] = A.m[3 * j + i];
	return B;
}
M3 axis_angle(V3 k, double th) {
	const double c = std::cos(th), s = std::sin(th), v = 1 - c;
	M3 R;
	R.m[0] = c + k.x * k.x * v;
	R.m[1] = k.x * k.y * v - k.z * s;
	R.m[2] = k.x * k.z * v + k.y * s;
	R.m[3] = k.y * k.x * v + k.z * s;
	R.m[4] = c + k.y * k.y * v;
	R.m[5] = k.y * k.z * v - k.x * s;
	R.m[6] = k.z * k.x * v - k.y * s;
	R.m[7] = k.z * k.y * v + k.x * s;
	R.m[8] = c + k.z * k.z * v;
	return R;
}
Mat toMat(const M3& A) {
	Mat B(3, 3);
	for (int i = 0; i < 9; i++) B.a[i] = A.m[i];
	return B;
}
Vec toVec(V3 v) { return {v.x, v.y, v.z}; }
V3 toV3(const Vec& v, int o = 0) { return {v[o], v[o + 1], v[o + 2]}; }

// SaiModel::orientationError
V3 orientationError(const M3& Rd, const M3& Rc) {
	V3 e;
	for (int k = 0; k < 3; k++) {
		V3 c{Rc.m[k], Rc.m[3 + k], Rc.m[6 + k]}, d{Rd.m[k], Rd.m[3 + k], Rd.m[6 + k]};
		e = e + cross(c, d);
	}
	return -0.5 * e;
}

// ------------------------------------------------------------------ robot model (sai-model stand-in)
struct ChainDesc {
	int n;
	std::vector<int> jtype;
	std::vector<V3> axis, t_fix, com;
	std::vector<M3> R_fix, inertia;
	Vec mass, q_lower, q_upper, effort;
};

struct Frame {
	int body;
	M3 R;
	V3 t;
};

struct OpSpace {
	Mat Lambda, Jbar, N;
};

struct Model {
	ChainDesc d;
	Vec q, dq;
	std::vector<M3> Rb;
	std::vector<V3> pb, ax;
	Mat M, Minv;
	explicit Model(const ChainDesc& desc) : d(desc), q(desc.n, 0.0), dq(desc.n, 0.0) { updateModel(); }
	int dof() const { return d.n; }
	void updateKinematics() {
		Rb.assign(d.n, M3());
		pb.assign(d.n, V3());
		ax.assign(d.n, V3());
		M3 R;
		V3 p;
		for (int i = 0; i < d.n; i++) {
			p = p + R * d.t_fix[i];
			R = R * d.R_fix[i];
			if (d.jtype[i] == 0)
				R = R * axis_angle(d.axis[i], q[i]);
			else
				p = p + q[i] * (R * d.axis[i]);
			Rb[i] = R;
			pb[i] = p;
			ax[i] = R * d.axis[i];
		}
	}
	Mat pointJacobian(int body, V3 x) const {
		Mat J(6, d.n);
		for (int i = 0; i <= body; i++) {
			if (d.jtype[i] == 0) {
				V3 v = cross(ax[i], x - pb[i]);
				J(0, i) = v.x;
				J(1, i) = v.y;
				J(2, i) = v.z;
				J(3, i) = ax[i].x;
				J(4, i) = ax[i].y;
				J(5, i) = ax[i].z;
			} else {
				J(0, i) = ax[i].x;
				J(1, i) = ax[i].y;
				J(2, i) = ax[i].z;
			}
		}
		return J;
	}
	// mass matrix by composite rigid bodies in link-local coordinates propagated to the base (RBDL-like recursion
	// written with explicit body Jacobians of the composite: M = sum_k m_k Jv^T Jv + Jw^T I Jw evaluated per body)
	void updateModel() {
		updateKinematics();
		const int n = d.n;
		M = Mat(n, n);
		for (int k = 0; k < n; k++) {
			const V3 c = pb[k] + Rb[k] * d.com[k];
			const Mat J = pointJacobian(k, c);
			const Mat Jv = block_rows(J, 0, 3), Jw = block_rows(J, 3, 3);
			const Mat Iw = toMat(Rb[k] * d.inertia[k] * Tr(Rb[k]));
			const Mat A = T(Jv) * Jv, B = T(Jw) * Iw * Jw;
			for (size_t i = 0; i < M.a.size(); i++) M.a[i] += d.mass[k] * A.a[i] + B.a[i];
		}
		Minv = inverse(M);
	}
	V3 position(const Frame& f, V3 p_in_link) const {
		const V3 local = f.t + f.R * p_in_link;
		return f.body < 0 ? local : pb[f.body] + Rb[f.body] * local;
	}
	M3 rotation(const Frame& f, const M3& R_in_link) const {
		const M3 Rl = f.R * R_in_link;
		return f.body < 0 ? Rl : Rb[f.body] * Rl;
	}
	Mat J(const Frame& f, V3 p_in_link) const {
		if (f.body < 0) return Mat(6, d.n);
		return pointJacobian(f.body, position(f, p_in_link));
	}
	OpSpace operationalSpaceMatrices(const Mat& Jt) const {
		OpSpace o;
		o.Lambda = llt_inverse(Jt * Minv * T(Jt));
		o.Jbar = Minv * T(Jt) * o.Lambda;
		o.N = Mat::identity(d.n) - o.Jbar * Jt;
		return o;
	}
	Vec gravityVector(V3 g) const {
		Vec tau(d.n, 0.0);
		for (int k = 0; k < d.n; k++) {
			const V3 c = pb[k] + Rb[k] * d.com[k];
			const Mat J = pointJacobian(k, c);
			for (int i = 0; i < d.n; i++) tau[i] += -d.mass[k] * (J(0, i) * g.x + J(1, i) * g.y + J(2, i) * g.z);
		}
		return tau;
	}
};

enum { FULL = 0, BIE = 1, IMPEDANCE = 2 };

Mat bie_inverse(const Model& robot, double thr) {
	Mat Mb = robot.M;
	for (int i = 0; i < robot.dof(); i++)
		if (Mb(i, i) < thr) Mb(i, i) = thr;
	return inverse(Mb);
}

// ------------------------------------------------------------------ POPC (POPCExplicitForceControl.cpp:5-96)
struct POPC {
	double dt;
	bool enabled = false;
	double PO = 0, Ecorr = 0, stored = 0, Rc = 1, vsum = 0;
	int counter = 50;
	std::deque<double> window;
	explicit POPC(double dt_) : dt(dt_) {}
	void reInitialize() {
		PO = Ecorr = stored = 0;
		window.clear();
		counter = 50;
		Rc = 1;
		vsum = 0;
	}
	V3 compute(V3 fd, V3 fs, V3 vcl, V3 vr, double kv, double kff) {
		if (!enabled) return vcl - kv * vr;
		const V3 Fcmd = kff * fd + Rc * vcl - kv * vr;
		const double vc2 = dot(vcl, vcl);
		const V3 fdiff = fs - fd;
		const double power = (dot(fdiff, vcl) - dot(Fcmd, vr)) * dt;
		PO += power;
		window.push_back(power);
		if (PO + stored + Ecorr > 0) {
			while (window.size() > 250) {
				if (PO + Ecorr + stored > window.front()) {
					if (window.front() > 0) PO -= window.front();
					window.pop_front();
				} else
					break;
			}
		}
		if (counter <= 0) {
			counter = 50;
			const double old = Rc;
			if (PO + stored + Ecorr < 0) {
				Rc = 1 + (PO + stored + Ecorr) / (vsum * dt);
				if (Rc > 1) Rc = 1;
				if (Rc < 0) Rc = 0;
			} else {
				Rc = (1 + (0.1 * 50 - 1) * Rc) / (double)(0.1 * 50);
			}
			Ecorr += (1 - old) * vsum * dt;
			vsum = 0;
		}
		counter--;
		vsum += vc2;
		return Rc * vcl - kv * vr;
	}
};

// ------------------------------------------------------------------ SingularityHandler (SingularityHandler.cpp:24-368)
struct SingularityHandler {
	Model* robot;
	Frame link;
	M3 cR;
	V3 ct;
	int task_rank, dof;
	int decoupling = BIE;
	double bie = 0.1;
	Vec q_upper, q_lower, tau_upper, tau_lower, type2_torque, q_prior, dq_prior, type2_dir;
	double kp1 = 50, kv1 = 14, kv2 = 5;
	double s_abs_tol = 1e-3, type1_tol = 0.5, type2_angle = 5 * M_PI / 180, perturb = 5, s_min = 6e-3, s_max = 6e-2;
	int buffer_size = 200;
	int c1 = 0, c2 = 0;
	bool enforce_type1 = false, handling = true;
	std::vector<int> types;
	std::deque<int> history;
	double alpha = 1;
	Mat N, U_ns, U_s, V_s, J_ns, J_s, L_ns, L_s, N_ns, L_js, J_post, L_ns_mod, L_s_mod, L_js_mod;
	Vec svd_s;
	SingularityHandler(Model* r, Frame l, M3 cR_, V3 ct_, int rank) : robot(r), link(l), cR(cR_), ct(ct_), task_rank(rank), dof(r->dof()) {
		const ChainDesc& d = r->d;
		q_upper = d.q_upper;
		q_lower = d.q_lower;
		tau_upper = d.effort;
		tau_lower = -1.0 * d.effort;
		type2_torque = 1e-2 * d.effort;	 // intended value
		q_prior = Vec(dof);
		for (int i = 0; i < dof; i++) q_prior[i] = 0.5 * (q_lower[i] + q_upper[i]);
		dq_prior = Vec(dof, 0.0);
		type2_dir = Vec(dof, 1.0);
		N = Mat(dof, dof);
		N_ns = Mat::identity(dof);
		U_ns = Mat(6, 1);
		U_s = Mat(6, 1);
		V_s = Mat(dof, 1);
		J_post = Mat(1, dof);
	}
	void updateTaskModel(const Mat& Jp, const Mat& N_prec) {
		const int r = task_rank, n = dof;
		SVD d = svd_thin(Jp);
		svd_s = d.s;
		const Mat& Minv = robot->Minv;
		if (d.s[0] < s_abs_tol) {
			alpha = 0;
			U_ns = Mat(r, 1);
			J_ns = Mat(r, n);
			L_ns = Mat(r, r);
			U_s = block_cols(d.U, 0, r);
			V_s = block_cols(d.V, 0, r);
			J_s = T(U_s) * Jp;
			L_s = pinv(J_s * Minv * T(J_s), 1e-15 * 8);
		} else {
			const int i0 = r > 1 ? 1 : 0;  // Appendix C6: rank 1 is always non-singular
			for (int i = i0; i < std::max(r, 1); i++) {
				const double c = d.s[i] / d.s[0];
				if (r > 1 && c < s_max) {
					alpha = std::min(std::max((c - s_min) / (s_max - s_min), 0.0), 1.0);
					U_ns = block_cols(d.U, 0, i);
					J_ns = T(U_ns) * Jp;
					OpSpace o = robot->operationalSpaceMatrices(J_ns);
					L_ns = o.Lambda;
					N_ns = o.N;
					U_s = block_cols(d.U, i, r - i);
					V_s = block_cols(d.V, i, r - i);
					J_s = T(U_s) * Jp;
					L_s = inverse(J_s * Minv * T(J_s));
					break;
				} else if (i == r - 1) {
					alpha = 1;
					U_ns = block_cols(d.U, 0, r);
					J_ns = T(U_ns) * Jp;
					OpSpace o = robot->operationalSpaceMatrices(J_ns);
					L_ns = o.Lambda;
					N_ns = o.N;
					U_s = Mat(r, 1);
					V_s = Mat(n, 1);
					J_s = Mat(r, n);
					L_s = Mat(r, r);
				}
			}
		}
		if (U_s.norm() == 0 || !handling) {
			N = N_ns;
			L_js = Mat(1, 1);
		} else if (U_ns.norm() == 0) {
			N = N_prec;
			L_js = Mat(1, 1);
		} else {
			J_post = T(V_s) * N_ns * N_prec;
			OpSpace o = robot->operationalSpaceMatrices(J_post);
			L_js = o.Lambda;
			N = o.N * N_ns;
		}
		if (decoupling == FULL) {
			L_ns_mod = L_ns;
			L_s_mod = L_s;
			L_js_mod = L_js;
		} else if (decoupling == IMPEDANCE) {
			L_ns_mod = Mat::identity(U_ns.c);
			L_s_mod = Mat::identity(U_s.c);
			L_js_mod = Mat::identity(V_s.c);
		} else {
			const Mat Mbi = bie_inverse(*robot, bie);
			L_ns_mod = U_ns.norm() != 0 ? inverse(J_ns * Mbi * T(J_ns)) : L_ns;
			L_s_mod = U_s.norm() != 0 ? inverse(J_s * Mbi * T(J_s)) : L_s;
			if (U_s.norm() != 0 && U_ns.norm() != 0 && handling)
				L_js_mod = inverse(J_post * Mbi * T(J_post));
			else
				L_js_mod = L_js;
		}
		classify();
	}
	void classify() {
		if (types.empty() || c2 > c1) {
			q_prior = robot->q;
			dq_prior = robot->dq;
		}
		if (U_s.norm() == 0) {
			types.clear();
			history.clear();
			c1 = c2 = 0;
			return;
		}
		const int k = U_s.c;
		types.assign(k, 0);
		const Vec q0 = robot->q;
		const V3 p0 = robot->position(link, ct);
		const M3 R0 = robot->rotation(link, cR);
		for (int i = 0; i < k; i++) {
			Vec qq = q0;
			for (int j = 0; j < dof; j++) qq[j] += perturb * V_s(j, i);
			robot->q = qq;
			robot->updateKinematics();
			const V3 dp = robot->position(link, ct) - p0;
			const V3 dphi = orientationError(robot->rotation(link, cR), R0);
			const double m = std::fabs(dp.x * U_s(0, i) + dp.y * U_s(1, i) + dp.z * U_s(2, i) + dphi.x * U_s(3, i) +
									   dphi.y * U_s(4, i) + dphi.z * U_s(5, i));
			types[i] = m > type1_tol ? 1 : 2;
			robot->q = q0;
			robot->updateKinematics();
		}
		bool any1 = std::find(types.begin(), types.end(), 1) != types.end();
		history.push_back(any1 ? 1 : 2);
		if (any1)
			c1++;
		else
			c2++;
		if ((int)history.size() > buffer_size) {
			if (history.front() == 1)
				c1--;
			else
				c2--;
			history.pop_front();
		}
	}
	Vec computeTorques(const Vec& f, const Vec& F) {
		auto ns_torque = [&]() { return T(J_ns) * ((L_ns_mod * (T(U_ns) * f)) + (T(U_ns) * F)); };
		if (types.empty()) return ns_torque();
		if (decoupling == IMPEDANCE) return T(J_ns) * ((T(U_ns) * f) + (T(U_ns) * F));
		if (U_ns.norm() == 0) return Vec(dof, 0.0);
		Vec tau_ns = ns_torque();
		if (!handling) return tau_ns;
		Vec js;
		const Vec &q = robot->q, &dq = robot->dq;
		if (c1 > c2 || enforce_type1) {
			Vec u = (-kp1) * (q - q_prior) - kv1 * dq;
			js = T(J_post) * (L_js_mod * (T(V_s) * u));
		} else {
			for (int i = 0; i < V_s.r; i++)
				if (V_s(i, 0) != 0) {
					if (std::fabs(q[i] - q_upper[i]) < type2_angle)
						type2_dir[i] = -1;
					else if (std::fabs(q[i] - q_lower[i]) < type2_angle)
						type2_dir[i] = 1;
				}
			Vec ff = f + F;
			const double nf = norm(ff);
			if (nf > 0) ff = (1.0 / nf) * ff;
			const double fTd = dot(ff, col(U_s, 0));
			Vec u(dof);
			for (int i = 0; i < dof; i++) u[i] = type2_dir[i] * std::fabs(fTd) * type2_torque[i];
			js = (T(J_post) * (T(V_s) * u)) + (T(J_post) * (L_js_mod * (T(V_s) * ((-kv2) * dq))));
		}
		Vec ts = T(J_s) * ((L_s_mod * (T(U_s) * f)) + (T(U_s) * F));
		for (int i = 0; i < dof; i++) {
			if (std::isnan(ts[i]))
				ts[i] = 0;
			else if (ts[i] > tau_upper[i])
				ts[i] = tau_upper[i];
			else if (ts[i] < tau_lower[i])
				ts[i] = tau_lower[i];
		}
		return tau_ns + alpha * ts + (1 - alpha) * js;
	}
};

// ------------------------------------------------------------------ tasks
struct Task {
	virtual ~Task() {}
	virtual void updateTaskModel(const Mat& N_prec) = 0;
	virtual Vec computeTorques() = 0;
	virtual Vec computeTorques(const Vec& tau_prec) = 0;
	virtual Mat getTaskAndPreviousNullspace() const = 0;
	virtual int type() const = 0;
};

// JointTask.cpp:14-356
struct JointTask : Task {
	Model* robot;
	Mat S;
	int k;
	double dt;
	int decoupling = BIE;
	double bie = 0.1;
	Vec kp, kv, ki, sat;
	bool use_sat = false;
	Vec goal_pos, goal_vel, goal_acc, integ;
	Mat N_prec, M_partial, M_mod, Jp, N, range;
	JointTask(Model* r, const Mat& S_, double dt_) : robot(r), S(S_), k(S_.r), dt(dt_) {
		kp = Vec(k, 50.0);
		kv = Vec(k, 14.0);
		ki = Vec(k, 0.0);
		sat = Vec(k, M_PI / 3);
		N_prec = Mat::identity(r->dof());
		range = Mat::identity(k);
		reInit();
	}
	void reInit() {
		goal_pos = S * robot->q;
		goal_vel = Vec(k, 0.0);
		goal_acc = Vec(k, 0.0);
		integ = Vec(k, 0.0);
	}
	int type() const override { return 2; }
	Mat getTaskAndPreviousNullspace() const override { return N * N_prec; }
	void updateTaskModel(const Mat& Np) override {
		const int n = robot->dof();
		N_prec = Np;
		Jp = S * N_prec;
		range = matrixRangeBasis(Jp);
		if (range.norm() == 0) {
			N = Mat::identity(n);
			return;
		}
		OpSpace o = robot->operationalSpaceMatrices(T(range) * Jp);
		M_partial = o.Lambda;
		N = o.N;
		if (decoupling == FULL)
			M_mod = M_partial;
		else if (decoupling == BIE)
			M_mod = inverse(T(range) * Jp * bie_inverse(*robot, bie) * T(Jp) * range);
		else
			M_mod = Mat::identity(range.c);
	}
	Vec computeTorques(const Vec& tau_prec) override {
		Vec tt = computeTorques();
		if (range.norm() == 0) return tt;
		Vec comp = T(Jp) * (range * (M_partial * (T(range) * (S * (robot->Minv * tau_prec)))));
		return tt - comp;
	}
	Vec computeTorques() override {
		Jp = S * N_prec;
		const Vec pos = S * robot->q, vel = S * robot->dq;
		if (range.norm() == 0) return Vec(robot->dof(), 0.0);
		Vec e = pos - goal_pos;
		integ = integ + dt * e;
		Vec t(k);
		if (use_sat) {
			for (int a = 0; a < k; a++) {
				const double kvi = kv[a] > 1e-6 ? 1.0 / kv[a] : 0.0;
				double vd = -kp[a] * kvi * e[a] - ki[a] * kvi * integ[a];
				if (vd > sat[a])
					vd = sat[a];
				else if (vd < -sat[a])
					vd = -sat[a];
				t[a] = -kv[a] * (vel[a] - vd);
			}
		} else {
			for (int a = 0; a < k; a++) t[a] = -kp[a] * e[a] - kv[a] * (vel[a] - goal_vel[a]) - ki[a] * integ[a];
		}
		Vec f = (M_partial * (T(range) * goal_acc)) + (M_mod * (T(range) * t));
		return T(Jp) * (range * f);
	}
};

// MotionForceTask.cpp:16-1001
struct MotionForceTask : Task {
	Model* robot;
	Frame link;
	M3 cR;
	V3 ct;
	bool in_compliant;
	double dt;
	Mat P;	// 6x6 partial task projection
	M3 Pt, Pr;
	int pos_range, ori_range;
	V3 kp_pos{100, 100, 100}, kv_pos{20, 20, 20}, ki_pos{0, 0, 0}, kp_ori{200, 200, 200}, kv_ori{28.3, 28.3, 28.3}, ki_ori{0, 0, 0};
	double kp_f = 0.7, kv_f = 10, ki_f = 1.3, kp_m = 0.7, kv_m = 10, ki_m = 1.3, kff_f = 0.95, kff_m = 0.95, max_f = 20, max_m = 10;
	bool use_sat = false;
	double lin_sat = 0.3, ang_sat = M_PI / 3;
	int fdim = 0, mdim = 0;
	V3 faxis, maxis;
	bool cl_force = false, cl_moment = false;
	M3 cs_R;
	V3 cs_t;
	POPC popc;
	std::unique_ptr<SingularityHandler> sh;
	V3 cur_pos, goal_pos, goal_v, goal_w, goal_a, goal_al, goal_force, goal_moment, sensed_f, sensed_m, Ip, Io, If, Im;
	M3 cur_ori, goal_ori;
	Mat J, Jp, N, N_prec;
	MotionForceTask(Model* r, Frame l, M3 cR_, V3 ct_, const Mat& P_, int pr, int orr, bool in_c, double dt_)
		: robot(r), link(l), cR(cR_), ct(ct_), in_compliant(in_c), dt(dt_), P(P_), pos_range(pr), ori_range(orr), popc(dt_) {
		for (int i = 0; i < 3; i++)
			for (int j = 0; j < 3; j++) {
				Pt.m[3 * i + j] = P(i, j);
				Pr.m[3 * i + j] = P(3 + i, 3 + j);
			}
		N_prec = Mat::identity(r->dof());
		sh.reset(new SingularityHandler(r, l, cR_, ct_, pr + orr));
		reInit();
	}
	void reInit() {
		cur_pos = robot->position(link, ct);
		goal_pos = cur_pos;
		cur_ori = robot->rotation(link, cR);
		goal_ori = cur_ori;
		goal_v = goal_w = goal_a = goal_al = goal_force = goal_moment = sensed_f = sensed_m = Ip = Io = If = Im = V3();
	}
	int type() const override { return 3; }
	Mat getTaskAndPreviousNullspace() const override { return N * N_prec; }
	M3 selRot() const { return in_compliant ? robot->rotation(link, cR) : M3(); }
	static M3 outer(V3 a) {
		M3 O;
		const double v[3] = {a.x, a.y, a.z};
		for (int i = 0; i < 3; i++)
			for (int j = 0; j < 3; j++) O.m[3 * i + j] = v[i] * v[j];
		return O;
	}
	static M3 sub(const M3& A, const M3& B) {
		M3 C;
		for (int i = 0; i < 9; i++) C.m[i] = A.m[i] - B.m[i];
		return C;
	}
	static M3 zero3() {
		M3 Z;
		for (int i = 0; i < 9; i++) Z.m[i] = 0;
		return Z;
	}
	M3 sigma(int dim, const M3& Psel, V3 axis) const {
		const M3 R = selRot();
		if (dim == 0) return zero3();
		if (dim == 1) return Psel * R * outer(axis) * Tr(R) * Tr(Psel);
		if (dim == 2) return Psel * sub(M3(), R * outer(axis) * Tr(R)) * Tr(Psel);
		return Psel;
	}
	M3 sigmaForce() const { return sigma(fdim, Pt, faxis); }
	M3 sigmaMoment() const { return sigma(mdim, Pr, maxis); }
	M3 sigmaPosition() const { return Pt * sub(M3(), sigmaForce()) * Tr(Pt); }
	M3 sigmaOrientation() const { return Pr * sub(M3(), sigmaMoment()) * Tr(Pr); }
	void updateSensed(V3 fs, V3 ms) {
		const M3 Rwc = robot->rotation(link, cR);
		const V3 f = cs_R * fs;
		const V3 m = cross(cs_t, f) + cs_R * ms;
		sensed_f = Rwc * f;
		sensed_m = Rwc * m;
	}
	void updateTaskModel(const Mat& Np) override {
		N_prec = Np;
		J = P * robot->J(link, ct);
		Jp = J * N_prec;
		sh->updateTaskModel(Jp, N_prec);
		N = sh->N;
	}
	Vec computeTorques(const Vec& tau_prec) override {
		// _Lambda is identically zero in the reference: the four products are still paid for
		Vec tt = computeTorques();
		Mat Lambda0(6, 6);
		Vec comp = T(Jp) * (Lambda0 * (J * (robot->Minv * tau_prec)));
		return tt - comp;
	}
	static V3 mulDiag(V3 k, V3 v) { return {k.x * v.x, k.y * v.y, k.z * v.z}; }
	static V3 pinvDiag(V3 k) { return {k.x > 1e-6 ? 1 / k.x : 0, k.y > 1e-6 ? 1 / k.y : 0, k.z > 1e-6 ? 1 / k.z : 0}; }
	static V3 clampNorm(V3 v, double mx) {
		const double n = std::sqrt(dot(v, v));
		return n > mx ? (mx / n) * v : v;
	}
	Vec computeTorques() override {
		const int n = robot->dof();
		J = P * robot->J(link, ct);
		Jp = J * N_prec;
		cur_pos = robot->position(link, ct);
		cur_ori = robot->rotation(link, cR);
		const Vec vw = J * robot->dq;
		const V3 v = toV3(vw, 0), w = toV3(vw, 3);
		if (pos_range + ori_range == 0) return Vec(n, 0.0);
		const M3 Sf = sigmaForce(), Sm = sigmaMoment(), Sp = sigmaPosition(), So = sigmaOrientation();
		const M3 Rsel = selRot();
		const V3 gf = Rsel * goal_force, gm = Rsel * goal_moment;
		V3 force_fb, moment_fb;
		if (cl_force) {
			If = If + dt * (Sf * (sensed_f - gf));
			V3 fb = Sf * ((-kp_f) * (sensed_f - gf) - ki_f * If);
			fb = clampNorm(fb, max_f);
			force_fb = popc.compute(Sf * gf, Sf * sensed_f, Sf * fb, Sf * v, kv_f, kff_f);
		} else {
			force_fb = Sf * ((-kv_f) * v);
		}
		if (cl_moment) {
			Im = Im + dt * (Sm * (sensed_m - gm));
			V3 mb = Sm * ((-kp_m) * (sensed_m - gm) - ki_m * Im);
			mb = clampNorm(mb, max_m);
			moment_fb = Sm * (mb - kv_m * w);
		} else {
			moment_fb = Sm * ((-kv_m) * w);
		}
		Ip = Ip + dt * (Sp * (cur_pos - goal_pos));
		V3 pos_force;
		if (use_sat) {
			const V3 kvi = pinvDiag(kv_pos);
			V3 vd = (-1.0) * mulDiag(kp_pos, mulDiag(kvi, Sp * (cur_pos - goal_pos))) - mulDiag(ki_pos, mulDiag(kvi, Ip));
			vd = clampNorm(vd, lin_sat);
			pos_force = Sp * (goal_a - mulDiag(kv_pos, v - vd));
		} else {
			pos_force = Sp * (goal_a - mulDiag(kp_pos, cur_pos - goal_pos) - mulDiag(kv_pos, v - goal_v) - mulDiag(ki_pos, Ip));
		}
		const V3 eo = So * orientationError(goal_ori, cur_ori);
		Io = Io + dt * eo;
		V3 ori_force;
		if (use_sat) {
			const V3 kvi = pinvDiag(kv_ori);
			V3 wd = (-1.0) * mulDiag(kp_ori, mulDiag(kvi, eo)) - mulDiag(ki_ori, mulDiag(kvi, Io));
			wd = clampNorm(wd, ang_sat);
			ori_force = So * (goal_al - mulDiag(kv_ori, w - wd));
		} else {
			ori_force = So * (goal_al - mulDiag(kp_ori, eo) - mulDiag(kv_ori, w - goal_w) - mulDiag(ki_ori, Io));
		}
		V3 ff_f = Sf * gf, ff_m = Sm * gm;
		if (cl_force) {
			ff_f = kff_f * ff_f;
			ff_m = kff_m * ff_m;
		}
		Vec f{pos_force.x, pos_force.y, pos_force.z, ori_force.x, ori_force.y, ori_force.z};
		const V3 Ff = force_fb + ff_f, Fm = moment_fb + ff_m;
		Vec F{Ff.x, Ff.y, Ff.z, Fm.x, Fm.y, Fm.z};
		return sh->computeTorques(f, F);
	}
};

// RobotController.cpp:68-118 for one robot
struct Controller {
	std::unique_ptr<Model> robot;
	std::vector<std::unique_ptr<Task>> tasks;
	bool gravity = false, saturation = false;
	V3 g{0, 0, -9.81};
	Vec cycle(bool use_prev) {
		const int n = robot->dof();
		Mat N_prec = Mat::identity(n);
		for (auto& t : tasks) {
			t->updateTaskModel(N_prec);
			N_prec = t->getTaskAndPreviousNullspace();
		}
		Vec tau(n, 0.0);
		for (auto& t : tasks) tau = tau + (use_prev ? t->computeTorques(tau) : t->computeTorques());
		if (saturation)
			for (int i = 0; i < n; i++) tau[i] = std::min(std::max(tau[i], -robot->d.effort[i]), robot->d.effort[i]);
		if (gravity) tau = tau + robot->gravityVector(g);
		return tau;
	}
};

struct Batch {
	ChainDesc desc;
	std::vector<Controller> ctl;
};

}  // namespace oref

using namespace oref;

extern "C" {

// model arrays are row-major per joint: axis[n][3], R_fix[n][9], t_fix[n][3], com[n][3], inertia[n][9]
void* oref_create(int n, const int* jtype, const double* axis, const double* R_fix, const double* t_fix, const double* mass,
				  const double* com, const double* inertia, const double* q_lower, const double* q_upper, const double* effort,
				  int n_robots) {
	Batch* b = new Batch();
	ChainDesc& d = b->desc;
	d.n = n;
	for (int i = 0; i < n; i++) {
		d.jtype.push_back(jtype[i]);
		d.axis.push_back({axis[3 * i], axis[3 * i + 1], axis[3 * i + 2]});
		d.t_fix.push_back({t_fix[3 * i], t_fix[3 * i + 1], t_fix[3 * i + 2]});
		d.com.push_back({com[3 * i], com[3 * i + 1], com[3 * i + 2]});
		M3 R, I;
		std::memcpy(R.m, R_fix + 9 * i, sizeof(R.m));
		std::memcpy(I.m, inertia + 9 * i, sizeof(I.m));
		d.R_fix.push_back(R);
		d.inertia.push_back(I);
		d.mass.push_back(mass[i]);
		d.q_lower.push_back(q_lower[i]);
		d.q_upper.push_back(q_upper[i]);
		d.effort.push_back(effort[i]);
	}
	b->ctl.resize(n_robots);
	for (auto& c : b->ctl) c.robot.reset(new Model(d));
	return b;
}
void oref_destroy(void* h) { delete (Batch*)h; }

// q, dq: [n_robots][n] row-major
void oref_set_state(void* h, const double* q, const double* dq) {
	Batch* b = (Batch*)h;
	const int n = b->desc.n;
	for (size_t i = 0; i < b->ctl.size(); i++) {
		Model& m = *b->ctl[i].robot;
		for (int j = 0; j < n; j++) {
			m.q[j] = q[i * n + j];
			m.dq[j] = dq[i * n + j];
		}
		m.updateModel();
	}
}

// P: 6x6 row-major projection; frame: body, R[9], t[3]; compliant R[9], t[3]
int oref_add_mft(void* h, int body, const double* link_R, const double* link_t, const double* cR, const double* ct, const double* P,
				 int pos_range, int ori_range, int in_compliant, double dt) {
	Batch* b = (Batch*)h;
	Frame f;
	f.body = body;
	std::memcpy(f.R.m, link_R, sizeof(f.R.m));
	f.t = {link_t[0], link_t[1], link_t[2]};
	M3 R;
	std::memcpy(R.m, cR, sizeof(R.m));
	Mat Pm(6, 6);
	for (int i = 0; i < 36; i++) Pm.a[i] = P[i];
	for (auto& c : b->ctl)
		c.tasks.emplace_back(new MotionForceTask(c.robot.get(), f, R, {ct[0], ct[1], ct[2]}, Pm, pos_range, ori_range, in_compliant != 0, dt));
	return (int)b->ctl[0].tasks.size() - 1;
}
int oref_add_jt(void* h, const double* S, int k, double dt) {
	Batch* b = (Batch*)h;
	const int n = b->desc.n;
	Mat Sm = S ? Mat(k, n) : Mat::identity(n);
	if (S)
		for (int i = 0; i < k * n; i++) Sm.a[i] = S[i];
	for (auto& c : b->ctl) c.tasks.emplace_back(new JointTask(c.robot.get(), Sm, dt));
	return (int)b->ctl[0].tasks.size() - 1;
}
void oref_set_decoupling(void* h, int task, int type, double bie) {
	Batch* b = (Batch*)h;
	for (auto& c : b->ctl) {
		Task* t = c.tasks[task].get();
		if (t->type() == 2) {
			((JointTask*)t)->decoupling = type;
			((JointTask*)t)->bie = bie;
		} else {
			((MotionForceTask*)t)->sh->decoupling = type;
			((MotionForceTask*)t)->sh->bie = bie;
		}
	}
}
void oref_set_options(void* h, int gravity, int saturation) {
	Batch* b = (Batch*)h;
	for (auto& c : b->ctl) {
		c.gravity = gravity != 0;
		c.saturation = saturation != 0;
	}
}
// joint goals: [n_robots][k] each
void oref_jt_set_goals(void* h, int task, const double* pos, const double* vel, const double* acc) {
	Batch* b = (Batch*)h;
	for (size_t i = 0; i < b->ctl.size(); i++) {
		JointTask* t = (JointTask*)b->ctl[i].tasks[task].get();
		for (int a = 0; a < t->k; a++) {
			t->goal_pos[a] = pos[i * t->k + a];
			t->goal_vel[a] = vel[i * t->k + a];
			t->goal_acc[a] = acc[i * t->k + a];
		}
	}
}
void oref_jt_set_gains(void* h, int task, const double* kp, const double* kv, const double* ki) {
	Batch* b = (Batch*)h;
	for (auto& c : b->ctl) {
		JointTask* t = (JointTask*)c.tasks[task].get();
		for (int a = 0; a < t->k; a++) {
			t->kp[a] = kp[a];
			t->kv[a] = kv[a];
			t->ki[a] = ki[a];
		}
	}
}
// mft goals per robot: pos[3], ori[9], v[3], w[3], a[3], al[3] -> 24 doubles, [n_robots][24]
void oref_mft_set_goals(void* h, int task, const double* g) {
	Batch* b = (Batch*)h;
	for (size_t i = 0; i < b->ctl.size(); i++) {
		MotionForceTask* t = (MotionForceTask*)b->ctl[i].tasks[task].get();
		const double* p = g + i * 24;
		t->goal_pos = {p[0], p[1], p[2]};
		std::memcpy(t->goal_ori.m, p + 3, 9 * sizeof(double));
		t->goal_v = {p[12], p[13], p[14]};
		t->goal_w = {p[15], p[16], p[17]};
		t->goal_a = {p[18], p[19], p[20]};
		t->goal_al = {p[21], p[22], p[23]};
	}
}
// current pose per robot: pos[3], ori[9] -> [n_robots][12]
void oref_mft_get_current(void* h, int task, double* out) {
	Batch* b = (Batch*)h;
	for (size_t i = 0; i < b->ctl.size(); i++) {
		MotionForceTask* t = (MotionForceTask*)b->ctl[i].tasks[task].get();
		double* p = out + i * 12;
		p[0] = t->cur_pos.x;
		p[1] = t->cur_pos.y;
		p[2] = t->cur_pos.z;
		std::memcpy(p + 3, t->cur_ori.m, 9 * sizeof(double));
	}
}
// force control set-up (examples/09-...cpp:175-183 and examples/07-...cpp:188-201)
void oref_mft_force_setup(void* h, int task, int fdim, const double* faxis, int mdim, const double* maxis, int cl_force, int cl_moment,
						  int passivity, const double* force_gains, const double* moment_gains) {
	Batch* b = (Batch*)h;
	for (auto& c : b->ctl) {
		MotionForceTask* t = (MotionForceTask*)c.tasks[task].get();
		auto unit = [](const double* a) {
			const double n = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
			return n > 0 ? V3{a[0] / n, a[1] / n, a[2] / n} : V3();
		};
		if (fdim != t->fdim || fdim == 1 || fdim == 2) {  // parametrizeForceMotionSpaces reset (:850-856)
			t->goal_pos = t->cur_pos;
			t->goal_v = t->goal_a = V3();
			t->Ip = t->If = V3();
		}
		t->fdim = fdim;
		t->faxis = unit(faxis);
		if (mdim != t->mdim || mdim == 1 || mdim == 2) {
			t->goal_ori = t->cur_ori;
			t->goal_w = t->goal_al = V3();
			t->Io = t->Im = V3();
		}
		t->mdim = mdim;
		t->maxis = unit(maxis);
		if (t->cl_force != (cl_force != 0)) t->Ip = t->If = V3();
		t->cl_force = cl_force != 0;
		if (t->cl_moment != (cl_moment != 0)) t->Io = t->Im = V3();
		t->cl_moment = cl_moment != 0;
		t->popc.enabled = passivity != 0;
		if (force_gains) {
			t->kp_f = force_gains[0];
			t->kv_f = force_gains[1];
			t->ki_f = force_gains[2];
		}
		if (moment_gains) {
			t->kp_m = moment_gains[0];
			t->kv_m = moment_gains[1];
			t->ki_m = moment_gains[2];
		}
	}
}
// goal force/moment [n_robots][6], sensed wrench in sensor frame [n_robots][6]
void oref_mft_set_force_goals(void* h, int task, const double* goal) {
	Batch* b = (Batch*)h;
	for (size_t i = 0; i < b->ctl.size(); i++) {
		MotionForceTask* t = (MotionForceTask*)b->ctl[i].tasks[task].get();
		t->goal_force = {goal[6 * i], goal[6 * i + 1], goal[6 * i + 2]};
		t->goal_moment = {goal[6 * i + 3], goal[6 * i + 4], goal[6 * i + 5]};
	}
}
void oref_mft_update_sensed(void* h, int task, const double* wrench) {
	Batch* b = (Batch*)h;
	for (size_t i = 0; i < b->ctl.size(); i++) {
		MotionForceTask* t = (MotionForceTask*)b->ctl[i].tasks[task].get();
		t->updateSensed({wrench[6 * i], wrench[6 * i + 1], wrench[6 * i + 2]}, {wrench[6 * i + 3], wrench[6 * i + 4], wrench[6 * i + 5]});
	}
}

// one control cycle for every robot: tau [n_robots][n]; n_threads >= 1 host threads over contiguous chunks
void oref_cycle(void* h, double* tau, int use_prev, int n_threads) {
	Batch* b = (Batch*)h;
	const int n = b->desc.n;
	const size_t N = b->ctl.size();
	auto work = [&](size_t lo, size_t hi) {
		for (size_t i = lo; i < hi; i++) {
			Vec t = b->ctl[i].cycle(use_prev != 0);
			for (int j = 0; j < n; j++) tau[i * n + j] = t[j];
		}
	};
	if (n_threads <= 1) {
		work(0, N);
		return;
	}
	std::vector<std::thread> th;
	const size_t chunk = (N + n_threads - 1) / n_threads;
	for (int k = 0; k < n_threads; k++) {
		const size_t lo = std::min(N, (size_t)k * chunk), hi = std::min(N, lo + chunk);
		if (lo < hi) th.emplace_back(work, lo, hi);
	}
	for (auto& t : th) t.join();
}

// one cycle including the model update (setQ/setDq/updateModel) -- what the reference's user loop does per cycle
void oref_step(void* h, const double* q, const double* dq, double* tau, int use_prev, int n_threads) {
	Batch* b = (Batch*)h;
	const int n = b->desc.n;
	const size_t N = b->ctl.size();
	auto work = [&](size_t lo, size_t hi) {
		for (size_t i = lo; i < hi; i++) {
			Model& m = *b->ctl[i].robot;
			for (int j = 0; j < n; j++) {
				m.q[j] = q[i * n + j];
				m.dq[j] = dq[i * n + j];
			}
			m.updateModel();
			Vec t = b->ctl[i].cycle(use_prev != 0);
			for (int j = 0; j < n; j++) tau[i * n + j] = t[j];
		}
	};
	if (n_threads <= 1) {
		work(0, N);
		return;
	}
	std::vector<std::thread> th;
	const size_t chunk = (N + n_threads - 1) / n_threads;
	for (int k = 0; k < n_threads; k++) {
		const size_t lo = std::min(N, (size_t)k * chunk), hi = std::min(N, lo + chunk);
		if (lo < hi) th.emplace_back(work, lo, hi);
	}
	for (auto& t : th) t.join();
}

int oref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
