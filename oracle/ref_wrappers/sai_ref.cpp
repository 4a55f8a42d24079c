// TEST INFRASTRUCTURE (oracle/_ref).  C entry points around the REFERENCE'S OWN control-law classes:
//   /root/reference/src/RobotController.cpp, src/tasks/{JointTask,MotionForceTask,SingularityHandler,JointLimitAvoidanceTask}.cpp,
//   src/helper_modules/{POPCExplicitForceControl,OTG_joints,OTG_6dof_cartesian,SaiPrimitivesCommonDefinitions}.cpp and the
//   vendored Ruckig -- compiled WHERE THEY LIE, UNMODIFIED, by oracle/Makefile against oracle/eigen_standin (stand-in for the
//   Eigen API those files use) and oracle/saimodel_standin (stand-in for the external sai-model), into oracle/_ref/libsai_ref.so.
// Nothing of the control law is restated in this file: it only builds the reference's objects, forwards setter/getter calls by
// name and loops the reference's updateControllerTaskModels() / computeControlTorques() over a batch of robots.
//
// This translation unit is compiled with -fno-access-control so that it can (i) read private members the parity tests compare
// (singularity types/counters, POPC state, integrators) and (ii) give SingularityHandler::_type_2_torque_vector a defined value:
// the reference builds it from _type_2_torque_ratio BEFORE that member is initialised (SingularityHandler.cpp:48 vs :69 --
// undefined behaviour).  sref_set_type2_ratio_fix(1) (default) rewrites the vector with the intended
// ratio 1e-2 after construction; with 0 the vector is left as the compiled code produced it.
//
// Used by tests/ (golden generation + parity) and by bench.py's CPU arms only.
#include <cstring>
#include <functional>
#include <map>
#include <sstream>
#include <thread>

#include "RobotController.h"

using namespace SaiPrimitives;

namespace {

using Args = std::vector<double>;

struct RefRobot {
	std::shared_ptr<SaiModel::SaiModel> model;
	std::vector<std::shared_ptr<TemplateTask>> tasks;
	std::unique_ptr<RobotController> ctl;
};
struct RefBatch {
	SaiModel::ChainDescription desc;
	std::vector<RefRobot> robots;
	std::string err;
};
bool g_type2_fix = true;

VectorXd vecx(const Args& a, size_t off, size_t n) {
	if (off + n > a.size()) throw std::invalid_argument("sai_ref: not enough arguments");
	VectorXd v(n);
	for (size_t i = 0; i < n; i++) v(i) = a[off + i];
	return v;
}
Vector3d v3(const Args& a, size_t off = 0) { return Vector3d(vecx(a, off, 3)); }
Matrix3d m3(const Args& a, size_t off = 0) {  // row-major
	if (off + 9 > a.size()) throw std::invalid_argument("sai_ref: not enough arguments");
	Matrix3d m;
	for (int r = 0; r < 3; r++)
		for (int c = 0; c < 3; c++) m(r, c) = a[off + 3 * r + c];
	return m;
}
template <class M>
Args flat(const M& m) {	 // row-major
	Args o;
	for (Eigen::Index r = 0; r < m.rows(); r++)
		for (Eigen::Index c = 0; c < m.cols(); c++) o.push_back(m(r, c));
	return o;
}
Args one(double x) { return Args{x}; }
Args none() { return Args{}; }
Args gains(const std::vector<PIDGains>& g) {
	Args o;
	for (const auto& x : g) {
		o.push_back(x.kp);
		o.push_back(x.kv);
		o.push_back(x.ki);
	}
	return o;
}
// scalar triple or three vectors of equal length
void split3(const Args& a, VectorXd& x, VectorXd& y, VectorXd& z) {
	const size_t k = a.size() / 3;
	x = vecx(a, 0, k);
	y = vecx(a, k, k);
	z = vecx(a, 2 * k, k);
}

using JFn = std::function<Args(JointTask&, const Args&)>;
using MFn = std::function<Args(MotionForceTask&, const Args&)>;
using CFn = std::function<Args(RefRobot&, const Args&)>;

const std::map<std::string, JFn>& joint_methods() {
	static const std::map<std::string, JFn> m = {
		{"setGoalPosition", [](JointTask& t, const Args& a) { t.setGoalPosition(vecx(a, 0, a.size())); return none(); }},
		{"setGoalVelocity", [](JointTask& t, const Args& a) { t.setGoalVelocity(vecx(a, 0, a.size())); return none(); }},
		{"setGoalAcceleration", [](JointTask& t, const Args& a) { t.setGoalAcceleration(vecx(a, 0, a.size())); return none(); }},
		{"setGains", [](JointTask& t, const Args& a) {
			 if (a.size() == 3) {
				 t.setGains(a[0], a[1], a[2]);
			 } else {
				 VectorXd x, y, z;
				 split3(a, x, y, z);
				 t.setGains(x, y, z);
			 }
			 return none();
		 }},
		{"setGainsUnsafe", [](JointTask& t, const Args& a) {
			 VectorXd x, y, z;
			 split3(a, x, y, z);
			 t.setGainsUnsafe(x, y, z);
			 return none();
		 }},
		{"getGains", [](JointTask& t, const Args&) { return gains(t.getGains()); }},
		{"setDynamicDecouplingType", [](JointTask& t, const Args& a) { t.setDynamicDecouplingType((DynamicDecouplingType)(int)a.at(0)); return none(); }},
		{"setBoundedInertiaEstimateThreshold", [](JointTask& t, const Args& a) { t.setBoundedInertiaEstimateThreshold(a.at(0)); return none(); }},
		{"getBoundedInertiaEstimateThreshold", [](JointTask& t, const Args&) { return one(t.getBoundedInertiaEstimateThreshold()); }},
		{"enableVelocitySaturation", [](JointTask& t, const Args& a) {
			 if (a.size() == 1)
				 t.enableVelocitySaturation(a[0]);
			 else
				 t.enableVelocitySaturation(vecx(a, 0, a.size()));
			 return none();
		 }},
		{"disableVelocitySaturation", [](JointTask& t, const Args&) { t.disableVelocitySaturation(); return none(); }},
		{"disableInternalOtg", [](JointTask& t, const Args&) { t.disableInternalOtg(); return none(); }},
		{"getInternalOtgEnabled", [](JointTask& t, const Args&) { return one(t.getInternalOtgEnabled()); }},
		{"enableInternalOtgAccelerationLimited", [](JointTask& t, const Args& a) {
			 if (a.size() == 2) {
				 t.enableInternalOtgAccelerationLimited(a[0], a[1]);
			 } else {
				 const size_t k = a.size() / 2;
				 t.enableInternalOtgAccelerationLimited(vecx(a, 0, k), vecx(a, k, k));
			 }
			 return none();
		 }},
		{"enableInternalOtgJerkLimited", [](JointTask& t, const Args& a) {
			 if (a.size() == 3) {
				 t.enableInternalOtgJerkLimited(a[0], a[1], a[2]);
			 } else {
				 VectorXd x, y, z;
				 split3(a, x, y, z);
				 t.enableInternalOtgJerkLimited(x, y, z);
			 }
			 return none();
		 }},
		{"reInitializeTask", [](JointTask& t, const Args&) { t.reInitializeTask(); return none(); }},
		{"resetIntegrators", [](JointTask& t, const Args&) { t.resetIntegrators(); return none(); }},
		{"goalPositionReached", [](JointTask& t, const Args& a) { return one(t.goalPositionReached(a.empty() ? 1e-2 : a[0])); }},
		{"getCurrentPosition", [](JointTask& t, const Args&) { return flat(t.getCurrentPosition()); }},
		{"getCurrentVelocity", [](JointTask& t, const Args&) { return flat(t.getCurrentVelocity()); }},
		{"getGoalPosition", [](JointTask& t, const Args&) { return flat(t.getGoalPosition()); }},
		{"getDesiredPosition", [](JointTask& t, const Args&) { return flat(t.getDesiredPosition()); }},
		{"getDesiredVelocity", [](JointTask& t, const Args&) { return flat(t.getDesiredVelocity()); }},
		{"getDesiredAcceleration", [](JointTask& t, const Args&) { return flat(t.getDesiredAcceleration()); }},
		{"getTaskNullspace", [](JointTask& t, const Args&) { return flat(t.getTaskNullspace()); }},
		{"getPreviousTasksNullspace", [](JointTask& t, const Args&) { return flat(t.getPreviousTasksNullspace()); }},
		{"getTaskAndPreviousNullspace", [](JointTask& t, const Args&) { return flat(t.getTaskAndPreviousNullspace()); }},
		{"computeTorques", [](JointTask& t, const Args& a) { return a.empty() ? flat(t.computeTorques()) : flat(t.computeTorques(vecx(a, 0, a.size()))); }},
		{"_integrated_position_error", [](JointTask& t, const Args&) { return flat(t._integrated_position_error); }},
		{"_M_partial", [](JointTask& t, const Args&) { return flat(t._M_partial); }},
		{"_M_partial_modified", [](JointTask& t, const Args&) { return flat(t._M_partial_modified); }},
		{"_current_task_range", [](JointTask& t, const Args&) { return flat(t._current_task_range); }},
	};
	return m;
}

const std::map<std::string, MFn>& mft_methods() {
	static const std::map<std::string, MFn> m = {
		{"setGoalPosition", [](MotionForceTask& t, const Args& a) { t.setGoalPosition(v3(a)); return none(); }},
		{"setGoalOrientation", [](MotionForceTask& t, const Args& a) { t.setGoalOrientation(m3(a)); return none(); }},
		{"setGoalLinearVelocity", [](MotionForceTask& t, const Args& a) { t.setGoalLinearVelocity(v3(a)); return none(); }},
		{"setGoalAngularVelocity", [](MotionForceTask& t, const Args& a) { t.setGoalAngularVelocity(v3(a)); return none(); }},
		{"setGoalLinearAcceleration", [](MotionForceTask& t, const Args& a) { t.setGoalLinearAcceleration(v3(a)); return none(); }},
		{"setGoalAngularAcceleration", [](MotionForceTask& t, const Args& a) { t.setGoalAngularAcceleration(v3(a)); return none(); }},
		{"setPosControlGains", [](MotionForceTask& t, const Args& a) {
			 if (a.size() == 3) {
				 t.setPosControlGains(a[0], a[1], a[2]);
			 } else {
				 VectorXd x, y, z;
				 split3(a, x, y, z);
				 t.setPosControlGains(x, y, z);
			 }
			 return none();
		 }},
		{"setOriControlGains", [](MotionForceTask& t, const Args& a) {
			 if (a.size() == 3) {
				 t.setOriControlGains(a[0], a[1], a[2]);
			 } else {
				 VectorXd x, y, z;
				 split3(a, x, y, z);
				 t.setOriControlGains(x, y, z);
			 }
			 return none();
		 }},
		{"getPosControlGains", [](MotionForceTask& t, const Args&) { return gains(t.getPosControlGains()); }},
		{"getOriControlGains", [](MotionForceTask& t, const Args&) { return gains(t.getOriControlGains()); }},
		{"setForceControlGains", [](MotionForceTask& t, const Args& a) { t.setForceControlGains(a.at(0), a.at(1), a.at(2)); return none(); }},
		{"setMomentControlGains", [](MotionForceTask& t, const Args& a) { t.setMomentControlGains(a.at(0), a.at(1), a.at(2)); return none(); }},
		{"setFeedforwardForceGain", [](MotionForceTask& t, const Args& a) { t.setFeedforwardForceGain(a.at(0)); return none(); }},
		{"setFeedforwardmomentGain", [](MotionForceTask& t, const Args& a) { t.setFeedforwardmomentGain(a.at(0)); return none(); }},
		{"setMaxForceControlFeedbackOutput", [](MotionForceTask& t, const Args& a) { t.setMaxForceControlFeedbackOutput(a.at(0)); return none(); }},
		{"setMaxMomentControlFeedbackOutput", [](MotionForceTask& t, const Args& a) { t.setMaxMomentControlFeedbackOutput(a.at(0)); return none(); }},
		{"setGoalForce", [](MotionForceTask& t, const Args& a) { t.setGoalForce(v3(a)); return none(); }},
		{"setGoalMoment", [](MotionForceTask& t, const Args& a) { t.setGoalMoment(v3(a)); return none(); }},
		{"getGoalForce", [](MotionForceTask& t, const Args&) { return flat(t.getGoalForce()); }},
		{"getGoalMoment", [](MotionForceTask& t, const Args&) { return flat(t.getGoalMoment()); }},
		{"enableVelocitySaturation", [](MotionForceTask& t, const Args& a) {
			 if (a.size() >= 2)
				 t.enableVelocitySaturation(a[0], a[1]);
			 else
				 t.enableVelocitySaturation();
			 return none();
		 }},
		{"disableVelocitySaturation", [](MotionForceTask& t, const Args&) { t.disableVelocitySaturation(); return none(); }},
		{"disableInternalOtg", [](MotionForceTask& t, const Args&) { t.disableInternalOtg(); return none(); }},
		{"getInternalOtgEnabled", [](MotionForceTask& t, const Args&) { return one(t.getInternalOtgEnabled()); }},
		{"enableInternalOtgAccelerationLimited", [](MotionForceTask& t, const Args& a) { t.enableInternalOtgAccelerationLimited(a.at(0), a.at(1), a.at(2), a.at(3)); return none(); }},
		{"enableInternalOtgJerkLimited", [](MotionForceTask& t, const Args& a) { t.enableInternalOtgJerkLimited(a.at(0), a.at(1), a.at(2), a.at(3), a.at(4), a.at(5)); return none(); }},
		// args: R (9, row-major), t (3) of the sensor frame in the link frame
		{"setForceSensorFrame", [](MotionForceTask& t, const Args& a) { t.setForceSensorFrame(t._link_name, Affine3d(m3(a, 0), v3(a, 9))); return none(); }},
		{"updateSensedForceAndMoment", [](MotionForceTask& t, const Args& a) { t.updateSensedForceAndMoment(v3(a, 0), v3(a, 3)); return none(); }},
		{"parametrizeForceMotionSpaces", [](MotionForceTask& t, const Args& a) { return one(a.size() >= 4 ? t.parametrizeForceMotionSpaces((int)a[0], v3(a, 1)) : t.parametrizeForceMotionSpaces((int)a.at(0))); }},
		{"parametrizeMomentRotMotionSpaces", [](MotionForceTask& t, const Args& a) { return one(a.size() >= 4 ? t.parametrizeMomentRotMotionSpaces((int)a[0], v3(a, 1)) : t.parametrizeMomentRotMotionSpaces((int)a.at(0))); }},
		{"setClosedLoopForceControl", [](MotionForceTask& t, const Args& a) { t.setClosedLoopForceControl(a.empty() || a[0] != 0); return none(); }},
		{"setClosedLoopMomentControl", [](MotionForceTask& t, const Args& a) { t.setClosedLoopMomentControl(a.empty() || a[0] != 0); return none(); }},
		{"enablePassivity", [](MotionForceTask& t, const Args&) { t.enablePassivity(); return none(); }},
		{"disablePassivity", [](MotionForceTask& t, const Args&) { t.disablePassivity(); return none(); }},
		{"setDynamicDecouplingType", [](MotionForceTask& t, const Args& a) { t.setDynamicDecouplingType((DynamicDecouplingType)(int)a.at(0)); return none(); }},
		{"setBoundedInertiaEstimateThreshold", [](MotionForceTask& t, const Args& a) { t.setBoundedInertiaEstimateThreshold(a.at(0)); return none(); }},
		{"handleAllSingularitiesAsType1", [](MotionForceTask& t, const Args& a) { t.handleAllSingularitiesAsType1(a.at(0) != 0); return none(); }},
		{"setType1Posture", [](MotionForceTask& t, const Args& a) { t.setType1Posture(vecx(a, 0, a.size())); return none(); }},
		{"enableSingularityHandling", [](MotionForceTask& t, const Args&) { t.enableSingularityHandling(); return none(); }},
		{"disableSingularityHandling", [](MotionForceTask& t, const Args&) { t.disableSingularityHandling(); return none(); }},
		{"setSingularityHandlingBounds", [](MotionForceTask& t, const Args& a) { t.setSingularityHandlingBounds(a.at(0), a.at(1)); return none(); }},
		{"setSingularityHandlingGains", [](MotionForceTask& t, const Args& a) { t.setSingularityHandlingGains(a.at(0), a.at(1), a.at(2)); return none(); }},
		{"reInitializeTask", [](MotionForceTask& t, const Args&) { t.reInitializeTask(); return none(); }},
		{"resetIntegrators", [](MotionForceTask& t, const Args&) { t.resetIntegrators(); return none(); }},
		{"resetIntegratorsLinear", [](MotionForceTask& t, const Args&) { t.resetIntegratorsLinear(); return none(); }},
		{"resetIntegratorsAngular", [](MotionForceTask& t, const Args&) { t.resetIntegratorsAngular(); return none(); }},
		{"goalPositionReached", [](MotionForceTask& t, const Args& a) { return one(t.goalPositionReached(a.at(0))); }},
		{"goalOrientationReached", [](MotionForceTask& t, const Args& a) { return one(t.goalOrientationReached(a.at(0))); }},
		{"getCurrentPosition", [](MotionForceTask& t, const Args&) { return flat(t.getCurrentPosition()); }},
		{"getCurrentOrientation", [](MotionForceTask& t, const Args&) { return flat(t.getCurrentOrientation()); }},
		{"getCurrentLinearVelocity", [](MotionForceTask& t, const Args&) { return flat(t.getCurrentLinearVelocity()); }},
		{"getCurrentAngularVelocity", [](MotionForceTask& t, const Args&) { return flat(t.getCurrentAngularVelocity()); }},
		{"getGoalPosition", [](MotionForceTask& t, const Args&) { return flat(t.getGoalPosition()); }},
		{"getGoalOrientation", [](MotionForceTask& t, const Args&) { return flat(t.getGoalOrientation()); }},
		{"getDesiredPosition", [](MotionForceTask& t, const Args&) { return flat(t.getDesiredPosition()); }},
		{"getDesiredOrientation", [](MotionForceTask& t, const Args&) { return flat(t.getDesiredOrientation()); }},
		{"getDesiredLinearVelocity", [](MotionForceTask& t, const Args&) { return flat(t.getDesiredLinearVelocity()); }},
		{"getDesiredAngularVelocity", [](MotionForceTask& t, const Args&) { return flat(t.getDesiredAngularVelocity()); }},
		{"getDesiredLinearAcceleration", [](MotionForceTask& t, const Args&) { return flat(t.getDesiredLinearAcceleration()); }},
		{"getDesiredAngularAcceleration", [](MotionForceTask& t, const Args&) { return flat(t.getDesiredAngularAcceleration()); }},
		{"getUnitMassForce", [](MotionForceTask& t, const Args&) { return flat(t.getUnitMassForce()); }},
		{"getPositionError", [](MotionForceTask& t, const Args&) { return flat(t.getPositionError()); }},
		{"getOrientationError", [](MotionForceTask& t, const Args&) { return flat(t.getOrientationError()); }},
		{"getSensedForceControlWorldFrame", [](MotionForceTask& t, const Args&) { return flat(t.getSensedForceControlWorldFrame()); }},
		{"getSensedMomentControlWorldFrame", [](MotionForceTask& t, const Args&) { return flat(t.getSensedMomentControlWorldFrame()); }},
		{"sigmaForce", [](MotionForceTask& t, const Args&) { return flat(t.sigmaForce()); }},
		{"sigmaPosition", [](MotionForceTask& t, const Args&) { return flat(t.sigmaPosition()); }},
		{"sigmaMoment", [](MotionForceTask& t, const Args&) { return flat(t.sigmaMoment()); }},
		{"sigmaOrientation", [](MotionForceTask& t, const Args&) { return flat(t.sigmaOrientation()); }},
		{"posSelectionProjector", [](MotionForceTask& t, const Args&) { return flat(t.posSelectionProjector()); }},
		{"oriSelectionProjector", [](MotionForceTask& t, const Args&) { return flat(t.oriSelectionProjector()); }},
		{"getTaskNullspace", [](MotionForceTask& t, const Args&) { return flat(t.getTaskNullspace()); }},
		{"getPreviousTasksNullspace", [](MotionForceTask& t, const Args&) { return flat(t.getPreviousTasksNullspace()); }},
		{"getTaskAndPreviousNullspace", [](MotionForceTask& t, const Args&) { return flat(t.getTaskAndPreviousNullspace()); }},
		{"computeTorques", [](MotionForceTask& t, const Args& a) { return a.empty() ? flat(t.computeTorques()) : flat(t.computeTorques(vecx(a, 0, a.size()))); }},
		// private state the parity tests compare
		{"_integrated_position_error", [](MotionForceTask& t, const Args&) { return flat(t._integrated_position_error); }},
		{"_integrated_orientation_error", [](MotionForceTask& t, const Args&) { return flat(t._integrated_orientation_error); }},
		{"_integrated_force_error", [](MotionForceTask& t, const Args&) { return flat(t._integrated_force_error); }},
		{"_integrated_moment_error", [](MotionForceTask& t, const Args&) { return flat(t._integrated_moment_error); }},
		{"_orientation_error", [](MotionForceTask& t, const Args&) { return flat(t._orientation_error); }},
		{"_current_task_range", [](MotionForceTask& t, const Args&) { return flat(t._current_task_range); }},
		{"_jacobian", [](MotionForceTask& t, const Args&) { return flat(t._jacobian); }},
		{"popc", [](MotionForceTask& t, const Args&) {
			 const auto& p = *t._POPC_force;
			 return Args{p._passivity_observer_value, p._E_correction, p._stored_energy_PO, (double)p._PO_counter, p._Rc, p._vcl_squared_sum, (double)p._PO_buffer_window.size(), p._is_enabled ? 1.0 : 0.0};
		 }},
		// [n_types, type_0.., counters 1 and 2, history length, alpha, n_singular_values, s...]
		{"singularity", [](MotionForceTask& t, const Args&) {
			 const auto& h = *t._singularity_handler;
			 Args o{(double)h._singularity_types.size()};
			 for (auto s : h._singularity_types) o.push_back((double)s);
			 o.push_back(h._type_1_counter);
			 o.push_back(h._type_2_counter);
			 o.push_back((double)h._singularity_history.size());
			 o.push_back(h._alpha);
			 o.push_back((double)h._svd_s.size());
			 for (Eigen::Index i = 0; i < h._svd_s.size(); i++) o.push_back(h._svd_s(i));
			 return o;
		 }},
		{"_q_prior", [](MotionForceTask& t, const Args&) { return flat(t._singularity_handler->_q_prior); }},
		{"_dq_prior", [](MotionForceTask& t, const Args&) { return flat(t._singularity_handler->_dq_prior); }},
		{"_type_2_direction", [](MotionForceTask& t, const Args&) { return flat(t._singularity_handler->_type_2_direction); }},
		{"_type_2_torque_vector", [](MotionForceTask& t, const Args&) { return flat(t._singularity_handler->_type_2_torque_vector); }},
		{"_svd_U", [](MotionForceTask& t, const Args&) { return flat(t._singularity_handler->_svd_U); }},
		{"_svd_V", [](MotionForceTask& t, const Args&) { return flat(t._singularity_handler->_svd_V); }},
		{"_task_range_s", [](MotionForceTask& t, const Args&) { return flat(t._singularity_handler->_task_range_s); }},
		{"_joint_task_range_s", [](MotionForceTask& t, const Args&) { return flat(t._singularity_handler->_joint_task_range_s); }},
	};
	return m;
}

const std::map<std::string, CFn>& controller_methods() {
	static const std::map<std::string, CFn> m = {
		{"enableGravityCompensation", [](RefRobot& r, const Args& a) { r.ctl->enableGravityCompensation(a.at(0) != 0); return none(); }},
		{"enableJointLimitAvoidance", [](RefRobot& r, const Args& a) { r.ctl->enableJointLimitAvoidance(a.at(0) != 0); return none(); }},
		{"enableTorqueSaturation", [](RefRobot& r, const Args& a) { r.ctl->enableTorqueSaturation(a.at(0) != 0); return none(); }},
		{"reinitializeTasks", [](RefRobot& r, const Args&) { r.ctl->reinitializeTasks(); return none(); }},
		{"updateControllerTaskModels", [](RefRobot& r, const Args&) { r.ctl->updateControllerTaskModels(); return none(); }},
		{"computeControlTorques", [](RefRobot& r, const Args&) { return flat(r.ctl->computeControlTorques()); }},
		// model-level queries (the sai-model stand-in)
		{"M", [](RefRobot& r, const Args&) { return flat(r.model->M()); }},
		{"MInv", [](RefRobot& r, const Args&) { return flat(r.model->MInv()); }},
		{"jointGravityVector", [](RefRobot& r, const Args&) { return flat(r.model->jointGravityVector()); }},
		{"q", [](RefRobot& r, const Args&) { return flat(r.model->q()); }},
		{"dq", [](RefRobot& r, const Args&) { return flat(r.model->dq()); }},
		{"jla_active_constraints", [](RefRobot& r, const Args&) { return one(r.ctl->_joint_limit_avoidance_task->_active_constraints); }},
		{"jla_limit_status", [](RefRobot& r, const Args&) {
			 Args o;
			 for (auto s : r.ctl->_joint_limit_avoidance_task->_limit_status) o.push_back((double)s);
			 return o;
		 }},
	};
	return m;
}

Vector3d rowv3(const double* p) { return Vector3d(p[0], p[1], p[2]); }
Matrix3d rowm3(const double* p) {
	Matrix3d m;
	for (int r = 0; r < 3; r++)
		for (int c = 0; c < 3; c++) m(r, c) = p[3 * r + c];
	return m;
}

// RobotController.cpp:68-118 driven exactly as the reference's user loop does; use_prev == 0 is the manual sum of
// examples/04-task_and_redundancy (tasks' computeTorques() without the previous-torque argument)
VectorXd cycle_one(RefRobot& r, bool use_prev) {
	r.ctl->updateControllerTaskModels();
	if (use_prev) return r.ctl->computeControlTorques();
	VectorXd tau = VectorXd::Zero(r.model->dof());
	for (auto& t : r.tasks) tau += t->computeTorques();
	return tau;
}

template <class F>
void parallel_for(size_t N, int n_threads, F work) {
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

}  // namespace

#define SREF_TRY try {
#define SREF_CATCH(b, ret)                    \
	}                                         \
	catch (const std::exception& e) {         \
		(b)->err = e.what();                  \
		return ret;                           \
	}

extern "C" {

void sref_set_type2_ratio_fix(int on) { g_type2_fix = on != 0; }
int sref_svd_oriented(void) {
#ifdef STANDIN_SVD_ORIENT
	return 1;
#else
	return 0;
#endif
}

// model arrays are row-major per joint (oracle/robots.py::Chain): axis[n][3], R_fix[n][9], t_fix[n][3], com[n][3], inertia[n][9];
// link frames: names separated by '\n', body[L], R[L][9], t[L][3]; base transform R[9], t[3] (may be null = identity)
void* sref_create(int n, const int* jtype, const double* axis, const double* R_fix, const double* t_fix, const double* mass, const double* com,
				  const double* inertia, const double* q_lower, const double* q_upper, const double* dq_max, const double* effort, const char* link_names,
				  const int* link_body, const double* link_R, const double* link_t, int n_links, const double* base_R, const double* base_t, int n_robots) {
	RefBatch* b = new RefBatch();
	auto& d = b->desc;
	d.n = n;
	for (int i = 0; i < n; i++) {
		d.jtype.push_back(jtype[i]);
		d.axis.push_back(rowv3(axis + 3 * i));
		d.t_fix.push_back(rowv3(t_fix + 3 * i));
		d.com.push_back(rowv3(com + 3 * i));
		d.R_fix.push_back(rowm3(R_fix + 9 * i));
		d.inertia.push_back(rowm3(inertia + 9 * i));
		d.mass.push_back(mass[i]);
		d.q_lower.push_back(q_lower[i]);
		d.q_upper.push_back(q_upper[i]);
		d.dq_max.push_back(dq_max[i]);
		d.effort.push_back(effort[i]);
	}
	std::istringstream names(link_names);
	std::string nm;
	for (int l = 0; l < n_links && std::getline(names, nm); l++) d.links[nm] = SaiModel::ChainDescription::LinkFrame{link_body[l], rowm3(link_R + 9 * l), rowv3(link_t + 3 * l)};
	b->robots.resize(n_robots);
	for (auto& r : b->robots) {
		r.model = std::make_shared<SaiModel::SaiModel>(d);
		if (base_R && base_t) r.model->setTRobotBase(Affine3d(rowm3(base_R), rowv3(base_t)));
	}
	return b;
}
void sref_destroy(void* h) { delete (RefBatch*)h; }
const char* sref_last_error(void* h) { return ((RefBatch*)h)->err.c_str(); }

// q, dq: [n_robots][n] row-major; SaiModel::setQ / setDq / updateModel as the reference's user loop does (examples/05-...cpp:143-145)
int sref_set_state(void* h, const double* q, const double* dq) {
	RefBatch* b = (RefBatch*)h;
	SREF_TRY
	const int n = b->desc.n;
	for (size_t i = 0; i < b->robots.size(); i++) {
		VectorXd qi(n), dqi(n);
		for (int j = 0; j < n; j++) {
			qi(j) = q[i * n + j];
			dqi(j) = dq[i * n + j];
		}
		b->robots[i].model->setQ(qi);
		b->robots[i].model->setDq(dqi);
		b->robots[i].model->updateModel();
	}
	return 0;
	SREF_CATCH(b, -1)
}

// MotionForceTask constructors (MotionForceTask.h:96-110).  partial == 0: full task; else direction lists dirs_t[n_t][3], dirs_r[n_r][3]
int sref_add_mft(void* h, const char* link, const double* cR, const double* ct, int partial, const double* dirs_t, int n_t, const double* dirs_r, int n_r,
				 int in_compliant, double dt, const char* name) {
	RefBatch* b = (RefBatch*)h;
	SREF_TRY
	const Affine3d frame(rowm3(cR), rowv3(ct));
	for (auto& r : b->robots) {
		std::shared_ptr<MotionForceTask> t;
		if (!partial) {
			t = std::make_shared<MotionForceTask>(r.model, link, frame, name, in_compliant != 0, dt);
		} else {
			std::vector<Vector3d> dt_, dr_;
			for (int i = 0; i < n_t; i++) dt_.push_back(rowv3(dirs_t + 3 * i));
			for (int i = 0; i < n_r; i++) dr_.push_back(rowv3(dirs_r + 3 * i));
			t = std::make_shared<MotionForceTask>(r.model, link, dt_, dr_, frame, name, in_compliant != 0, dt);
		}
		if (g_type2_fix) {	// Appendix C3, see the header comment
			auto& sh = *t->_singularity_handler;
			const auto limits = r.model->jointLimits();
			for (size_t i = 0; i < limits.size(); i++) sh._type_2_torque_vector(i) = 1e-2 * limits[i].effort;
			sh._type_2_torque_ratio = 1e-2;
		}
		r.tasks.push_back(t);
	}
	return (int)b->robots[0].tasks.size() - 1;
	SREF_CATCH(b, -1)
}
// JointTask constructors (JointTask.h:56-75).  S == null: full joint task
int sref_add_jt(void* h, const double* S, int k, double dt, const char* name) {
	RefBatch* b = (RefBatch*)h;
	SREF_TRY
	const int n = b->desc.n;
	for (auto& r : b->robots) {
		if (!S) {
			r.tasks.push_back(std::make_shared<JointTask>(r.model, name, dt));
		} else {
			MatrixXd Sm(k, n);
			for (int i = 0; i < k; i++)
				for (int j = 0; j < n; j++) Sm(i, j) = S[i * n + j];
			r.tasks.push_back(std::make_shared<JointTask>(r.model, Sm, name, dt));
		}
	}
	return (int)b->robots[0].tasks.size() - 1;
	SREF_CATCH(b, -1)
}
// RobotController(robot, tasks) (RobotController.cpp:8-66)
int sref_finalize(void* h) {
	RefBatch* b = (RefBatch*)h;
	SREF_TRY
	for (auto& r : b->robots) r.ctl = std::make_unique<RobotController>(r.model, r.tasks);
	return 0;
	SREF_CATCH(b, -1)
}

// forwards one call by method name.  task >= 0: a task of robot `robot`; task == -1: the robot's controller / model.
// robot == -1: every robot (arguments shared, result of the last robot returned).  Returns the number of doubles written to out,
// -1 on an exception thrown by the reference (message in sref_last_error), -2 for an unknown method, -3 when out is too small.
int sref_call(void* h, int task, int robot, const char* method, const double* in, int n_in, double* out, int out_cap) {
	RefBatch* b = (RefBatch*)h;
	SREF_TRY
	const Args a(in, in + n_in);
	Args res;
	const size_t lo = robot < 0 ? 0 : (size_t)robot, hi = robot < 0 ? b->robots.size() : (size_t)robot + 1;
	if (hi > b->robots.size()) throw std::invalid_argument("sai_ref: robot index out of range");
	for (size_t i = lo; i < hi; i++) {
		RefRobot& r = b->robots[i];
		if (task < 0) {
			auto it = controller_methods().find(method);
			if (it == controller_methods().end()) return -2;
			res = it->second(r, a);
		} else {
			TemplateTask* t = r.tasks.at(task).get();
			if (t->getTaskType() == TaskType::JOINT_TASK) {
				auto it = joint_methods().find(method);
				if (it == joint_methods().end()) return -2;
				res = it->second(*static_cast<JointTask*>(t), a);
			} else {
				auto it = mft_methods().find(method);
				if (it == mft_methods().end()) return -2;
				res = it->second(*static_cast<MotionForceTask*>(t), a);
			}
		}
	}
	if ((int)res.size() > out_cap) return -3;
	for (size_t i = 0; i < res.size(); i++) out[i] = res[i];
	return (int)res.size();
	SREF_CATCH(b, -1)
}

// bulk goal setters for the timed CPU arms (bench.py): g [n_robots][24] = position 3, orientation 9 (row-major), linear and
// angular velocity 3 + 3, linear and angular acceleration 3 + 3; joint goal positions [n_robots][k]
int sref_mft_set_goals(void* h, int task, const double* g) {
	RefBatch* b = (RefBatch*)h;
	SREF_TRY
	for (size_t i = 0; i < b->robots.size(); i++) {
		auto* t = static_cast<MotionForceTask*>(b->robots[i].tasks.at(task).get());
		const double* p = g + 24 * i;
		t->setGoalPosition(rowv3(p));
		t->setGoalOrientation(rowm3(p + 3));
		t->setGoalLinearVelocity(rowv3(p + 12));
		t->setGoalAngularVelocity(rowv3(p + 15));
		t->setGoalLinearAcceleration(rowv3(p + 18));
		t->setGoalAngularAcceleration(rowv3(p + 21));
	}
	return 0;
	SREF_CATCH(b, -1)
}
int sref_jt_set_goal_positions(void* h, int task, const double* pos) {
	RefBatch* b = (RefBatch*)h;
	SREF_TRY
	for (size_t i = 0; i < b->robots.size(); i++) {
		auto* t = static_cast<JointTask*>(b->robots[i].tasks.at(task).get());
		const int k = t->getTaskDof();
		VectorXd v(k);
		for (int j = 0; j < k; j++) v(j) = pos[(size_t)k * i + j];
		t->setGoalPosition(v);
	}
	return 0;
	SREF_CATCH(b, -1)
}

// one control cycle for every robot: tau [n_robots][n]
int sref_cycle(void* h, double* tau, int use_prev, int n_threads) {
	RefBatch* b = (RefBatch*)h;
	SREF_TRY
	const int n = b->desc.n;
	std::string err;
	parallel_for(b->robots.size(), n_threads, [&](size_t lo, size_t hi) {
		try {
			for (size_t i = lo; i < hi; i++) {
				const VectorXd t = cycle_one(b->robots[i], use_prev != 0);
				for (int j = 0; j < n; j++) tau[i * n + j] = t(j);
			}
		} catch (const std::exception& e) {
			err = e.what();
		}
	});
	if (!err.empty()) throw std::runtime_error(err);
	return 0;
	SREF_CATCH(b, -1)
}
// model update + cycle: what the reference's user loop does per control cycle
int sref_step(void* h, const double* q, const double* dq, double* tau, int use_prev, int n_threads) {
	RefBatch* b = (RefBatch*)h;
	SREF_TRY
	const int n = b->desc.n;
	std::string err;
	parallel_for(b->robots.size(), n_threads, [&](size_t lo, size_t hi) {
		try {
			for (size_t i = lo; i < hi; i++) {
				VectorXd qi(n), dqi(n);
				for (int j = 0; j < n; j++) {
					qi(j) = q[i * n + j];
					dqi(j) = dq[i * n + j];
				}
				auto& r = b->robots[i];
				r.model->setQ(qi);
				r.model->setDq(dqi);
				r.model->updateModel();
				const VectorXd t = cycle_one(r, use_prev != 0);
				for (int j = 0; j < n; j++) tau[i * n + j] = t(j);
			}
		} catch (const std::exception& e) {
			err = e.what();
		}
	});
	if (!err.empty()) throw std::runtime_error(err);
	return 0;
	SREF_CATCH(b, -1)
}
int sref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
