// TEST INFRASTRUCTURE (oracle).  NOT sai-model.
//
// The reference keeps its kinematics/dynamics in the un-vendored dependency sai-model (github manips-sai-org/sai-model,
// taken @master by the reference's CI, .github/actions/build-repo/action.yml:20-23; wraps RBDL).  It is absent from
// /root/reference and from this container.  This header declares a class with the NAMES AND SIGNATURES the reference's
// sources call (call sites: SURVEY.md section 8c) so that those sources compile where they lie, unmodified
// (oracle/Makefile -> oracle/_ref/libsai_ref.so), and implements each call with the standard algorithm behind it
// (SURVEY.md Appendix B):
//   updateKinematics / updateModel   forward kinematics of a serial chain; mass matrix from its definition
//                                    sum_k m_k Jv^T Jv + Jw^T I_k Jw; MInv = M.inverse() (LU, partial pivoting)
//   J / JWorldFrame                  6 x n point Jacobian, linear rows first (row order: MotionForceTask.cpp:293-298)
//   position / rotation [InWorld]    point position, frame orientation; transformInWorld
//   operationalSpaceMatrices         Lambda = (J MInv J^T).llt().solve(I), Jbar = MInv J^T Lambda, N = I - Jbar J
//   matrixRangeBasis                 thin SVD range basis, relative tolerance 1e-3, Identity when full row rank
//   orientationError                 -1/2 sum_i Rc[:,i] x Rd[:,i]
//   computePseudoInverse             SVD Moore-Penrose
//   jointGravityVector, jointLimits
// The CONTROL LAW is the reference's own compiled code; the MODEL ARITHMETIC under it is this restatement (the same
// algorithms as oracle/sai_model.py, "arithmetic unpinned by necessity").  The robot is described by the merged serial
// chain of oracle/robots.py (fixed-joint bodies merged into their parent like RBDL's URDF reader does).
#pragma once
#include <Eigen/Dense>

#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

// sai-model's public header opens both namespaces; the reference's headers rely on it (JointTask.h uses vector<>,
// shared_ptr<> unqualified with only `using namespace Eigen` of its own).
using namespace std;
using namespace Eigen;

namespace SaiModel {

struct JointLimit {
	std::string joint_name;
	int joint_index;
	double position_lower, position_upper, velocity, effort;
	JointLimit(const std::string& name, int index, double lower, double upper, double vel, double f)
		: joint_name(name), joint_index(index), position_lower(lower), position_upper(upper), velocity(vel), effort(f) {}
};

struct OpSpaceMatrices {
	MatrixXd J, Lambda, Jbar, N;
	OpSpaceMatrices(const MatrixXd& J_, const MatrixXd& L_, const MatrixXd& Jb_, const MatrixXd& N_) : J(J_), Lambda(L_), Jbar(Jb_), N(N_) {}
};

// merged serial chain: body i is moved by joint i (oracle/robots.py::Chain)
struct ChainDescription {
	int n = 0;
	std::vector<int> jtype;	 // 0 revolute, 1 prismatic
	std::vector<Vector3d> axis, t_fix, com;
	std::vector<Matrix3d> R_fix, inertia;
	std::vector<double> mass, q_lower, q_upper, dq_max, effort;
	struct LinkFrame {
		int body;
		Matrix3d R;
		Vector3d t;
	};
	std::map<std::string, LinkFrame> links;
};

inline Matrix3d axisAngle(const Vector3d& k, double th) {
	Matrix3d K = Matrix3d::Zero();
	K(0, 1) = -k(2);
	K(0, 2) = k(1);
	K(1, 0) = k(2);
	K(1, 2) = -k(0);
	K(2, 0) = -k(1);
	K(2, 1) = k(0);
	return Matrix3d::Identity() + std::sin(th) * K + (1.0 - std::cos(th)) * (K * K);
}

MatrixXd matrixRangeBasis(const MatrixXd& matrix, const double& tolerance = 1e-3);
Vector3d orientationError(const Matrix3d& desired_orientation, const Matrix3d& current_orientation);
MatrixXd computePseudoInverse(const MatrixXd& matrix, const double& tolerance = 1e-6);

class SaiModel {
public:
	explicit SaiModel(const ChainDescription& chain) : _c(chain), _dof(chain.n) {
		_q = VectorXd::Zero(_dof);
		_dq = VectorXd::Zero(_dof);
		_ddq = VectorXd::Zero(_dof);
		_T_world_robot = Affine3d::Identity();
		_world_gravity = Vector3d(0, 0, -9.81);
		for (int i = 0; i < _dof; i++) _joint_limits.push_back(JointLimit("joint" + std::to_string(i), i, chain.q_lower[i], chain.q_upper[i], chain.dq_max[i], chain.effort[i]));
		updateModel();
	}

	const int& dof() const { return _dof; }
	const VectorXd& q() const { return _q; }
	const VectorXd& dq() const { return _dq; }
	void setQ(const VectorXd& q) {
		if (q.size() != _dof) throw std::invalid_argument("q size inconsistent in SaiModel::setQ");
		_q = q;
	}
	void setDq(const VectorXd& dq) {
		if (dq.size() != _dof) throw std::invalid_argument("dq size inconsistent in SaiModel::setDq");
		_dq = dq;
	}
	const MatrixXd& M() const { return _M; }
	const MatrixXd& MInv() const { return _M_inv; }
	const Affine3d& TRobotBase() const { return _T_world_robot; }
	void setTRobotBase(const Affine3d& T) { _T_world_robot = T; }
	const std::vector<JointLimit>& jointLimits() const { return _joint_limits; }

	void updateKinematics() {
		_Rb.assign(_dof, Matrix3d::Identity());
		_pb.assign(_dof, Vector3d::Zero());
		_ax.assign(_dof, Vector3d::Zero());
		Matrix3d R = Matrix3d::Identity();
		Vector3d p = Vector3d::Zero();
		for (int i = 0; i < _dof; i++) {
			p = p + R * _c.t_fix[i];
			R = R * _c.R_fix[i];
			if (_c.jtype[i] == 0)
				R = R * axisAngle(_c.axis[i], _q(i));
			else
				p = p + _q(i) * (R * _c.axis[i]);
			_Rb[i] = R;
			_pb[i] = p;
			_ax[i] = R * _c.axis[i];
		}
	}
	void updateModel() {
		updateKinematics();
		_M = MatrixXd::Zero(_dof, _dof);
		for (int k = 0; k < _dof; k++) {
			const Vector3d pc = _pb[k] + _Rb[k] * _c.com[k];
			const MatrixXd Jk = bodyPointJacobian(k, pc);
			const MatrixXd Jv = Jk.block(0, 0, 3, _dof), Jw = Jk.block(3, 0, 3, _dof);
			const Matrix3d Iw = _Rb[k] * _c.inertia[k] * _Rb[k].transpose();
			_M += _c.mass[k] * (Jv.transpose() * Jv) + Jw.transpose() * Iw * Jw;
		}
		_M = 0.5 * (_M + _M.transpose());
		_M_inv = _M.inverse();
	}

	Vector3d position(const std::string& link_name, const Vector3d& pos_in_link = Vector3d::Zero()) const {
		const auto& f = link(link_name);
		const Vector3d local = f.t + f.R * pos_in_link;
		if (f.body < 0) return local;
		return _pb[f.body] + _Rb[f.body] * local;
	}
	Matrix3d rotation(const std::string& link_name, const Matrix3d& rot_in_link = Matrix3d::Identity()) const {
		const auto& f = link(link_name);
		const Matrix3d Rl = f.R * rot_in_link;
		if (f.body < 0) return Rl;
		return _Rb[f.body] * Rl;
	}
	Vector3d positionInWorld(const std::string& link_name, const Vector3d& pos_in_link = Vector3d::Zero()) const {
		return _T_world_robot * position(link_name, pos_in_link);
	}
	Matrix3d rotationInWorld(const std::string& link_name, const Matrix3d& rot_in_link = Matrix3d::Identity()) const {
		return _T_world_robot.linear() * rotation(link_name, rot_in_link);
	}
	Affine3d transformInWorld(const std::string& link_name, const Affine3d& T_in_link = Affine3d::Identity()) const {
		return Affine3d(rotationInWorld(link_name, T_in_link.linear()), positionInWorld(link_name, T_in_link.translation()));
	}
	MatrixXd J(const std::string& link_name, const Vector3d& pos_in_link = Vector3d::Zero()) const {
		const auto& f = link(link_name);
		if (f.body < 0) return MatrixXd::Zero(6, _dof);
		return bodyPointJacobian(f.body, position(link_name, pos_in_link));
	}
	MatrixXd JWorldFrame(const std::string& link_name, const Vector3d& pos_in_link = Vector3d::Zero()) const {
		MatrixXd Jb = J(link_name, pos_in_link);
		MatrixXd Jw(6, _dof);
		Jw.block(0, 0, 3, _dof) = _T_world_robot.linear() * Jb.block(0, 0, 3, _dof);
		Jw.block(3, 0, 3, _dof) = _T_world_robot.linear() * Jb.block(3, 0, 3, _dof);
		return Jw;
	}
	OpSpaceMatrices operationalSpaceMatrices(const MatrixXd& task_jacobian) const {
		if (task_jacobian.cols() != _dof) throw std::invalid_argument("Jacobian size inconsistent with DOF of robot model in SaiModel::operationalSpaceMatrices");
		const MatrixXd inv_inertia = task_jacobian * _M_inv * task_jacobian.transpose();
		const MatrixXd Lambda = inv_inertia.llt().solve(MatrixXd::Identity(task_jacobian.rows(), task_jacobian.rows()));
		const MatrixXd Jbar = _M_inv * task_jacobian.transpose() * Lambda;
		const MatrixXd N = MatrixXd::Identity(_dof, _dof) - Jbar * task_jacobian;
		return OpSpaceMatrices(task_jacobian, Lambda, Jbar, N);
	}
	MatrixXd nullspaceMatrix(const MatrixXd& task_jacobian) const { return operationalSpaceMatrices(task_jacobian).N; }
	VectorXd jointGravityVector() const {
		const Vector3d g_base = _T_world_robot.linear().transpose() * _world_gravity;
		VectorXd tau = VectorXd::Zero(_dof);
		for (int k = 0; k < _dof; k++) {
			const Vector3d pc = _pb[k] + _Rb[k] * _c.com[k];
			const MatrixXd Jk = bodyPointJacobian(k, pc);
			tau += Jk.block(0, 0, 3, _dof).transpose() * (-_c.mass[k] * g_base);
		}
		return tau;
	}

private:
	const ChainDescription::LinkFrame& link(const std::string& name) const {
		auto it = _c.links.find(name);
		if (it == _c.links.end()) throw std::invalid_argument("link [" + name + "] does not exist");
		return it->second;
	}
	MatrixXd bodyPointJacobian(int body, const Vector3d& p_base) const {
		MatrixXd Jm = MatrixXd::Zero(6, _dof);
		for (int i = 0; i <= body; i++) {
			if (_c.jtype[i] == 0) {
				const Vector3d v = _ax[i].cross(p_base - _pb[i]);
				for (int r = 0; r < 3; r++) {
					Jm(r, i) = v(r);
					Jm(3 + r, i) = _ax[i](r);
				}
			} else {
				for (int r = 0; r < 3; r++) Jm(r, i) = _ax[i](r);
			}
		}
		return Jm;
	}

	ChainDescription _c;
	int _dof;
	VectorXd _q, _dq, _ddq;
	MatrixXd _M, _M_inv;
	Affine3d _T_world_robot;
	Vector3d _world_gravity;
	std::vector<JointLimit> _joint_limits;
	std::vector<Matrix3d> _Rb;
	std::vector<Vector3d> _pb, _ax;
};

inline MatrixXd matrixRangeBasis(const MatrixXd& matrix, const double& tolerance) {
	const int range_size = matrix.rows();
	if (matrix.norm() < tolerance) return MatrixXd::Zero(range_size, 1);
	JacobiSVD<MatrixXd> svd(matrix, ComputeThinU | ComputeThinV);
	const double sigma_0 = svd.singularValues()(0);
	if (sigma_0 < tolerance) return MatrixXd::Zero(range_size, 1);
	int task_dof = std::min(matrix.rows(), matrix.cols());
	for (int i = svd.singularValues().size() - 1; i > 0; i--) {
		if (svd.singularValues()(i) / sigma_0 < tolerance)
			task_dof -= 1;
		else
			break;
	}
	if (task_dof == range_size) return MatrixXd::Identity(range_size, range_size);
	return svd.matrixU().leftCols(task_dof);
}

inline Vector3d orientationError(const Matrix3d& desired_orientation, const Matrix3d& current_orientation) {
	const Matrix3d Q1 = desired_orientation * desired_orientation.transpose() - Matrix3d::Identity();
	const Matrix3d Q2 = current_orientation * current_orientation.transpose() - Matrix3d::Identity();
	if (Q1.norm() > 0.0001 || Q2.norm() > 0.0001) throw std::invalid_argument("Invalid rotation matrices in SaiModel::orientationError");
	Vector3d e = Vector3d::Zero();
	for (int i = 0; i < 3; i++) e += current_orientation.col(i).cross(desired_orientation.col(i));
	return -0.5 * e;
}

inline MatrixXd computePseudoInverse(const MatrixXd& matrix, const double& tolerance) {
	JacobiSVD<MatrixXd> svd(matrix, ComputeThinU | ComputeThinV);
	MatrixXd out = MatrixXd::Zero(matrix.cols(), matrix.rows());
	for (int i = 0; i < svd.singularValues().size(); i++) {
		const double s = svd.singularValues()(i);
		if (s > tolerance) out += (svd.matrixV().col(i) * svd.matrixU().col(i).transpose()) / s;
	}
	return out;
}

}  // namespace SaiModel
