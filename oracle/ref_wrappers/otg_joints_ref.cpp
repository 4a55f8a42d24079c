// TEST INFRASTRUCTURE (oracle/_ref), groundwork for SURVEY.md row f-4 (batched OTG; not built in the product yet).
// The trajectory generator is the REFERENCE'S vendored Ruckig (/root/reference/ruckig, compiled where it lies by
// oracle/Makefile).  Around it sits a restatement of the reference's thin wrapper
// src/helper_modules/OTG_joints.cpp:17-150 (that file itself needs Eigen's dynamic vectors): phase synchronisation (:24),
// disableJerkLimits (:89-92), setGoalPositionAndVelocity (:99-116), update (:118-150).
#include <cmath>
#include <limits>
#include <vector>

#include <ruckig/ruckig.hpp>

namespace {
struct OtgJoints {
	int dim;
	ruckig::Ruckig<ruckig::DynamicDOFs> otg;
	ruckig::InputParameter<ruckig::DynamicDOFs> in;
	ruckig::OutputParameter<ruckig::DynamicDOFs> out;
	bool goal_reached = true;
	std::vector<double> goal;
	OtgJoints(int n, double dt) : dim(n), otg(n, dt), in(n), out(n), goal(n, 0.0) { in.synchronization = ruckig::Synchronization::Phase; }
	void set_goal(const double* p, const double* v) {  // :99-116 (isApprox with Eigen's default precision 1e-12)
		auto approx = [&](const std::vector<double>& a, const double* b) {
			double d2 = 0, na = 0, nb = 0;
			for (int i = 0; i < dim; i++) {
				d2 += (a[i] - b[i]) * (a[i] - b[i]);
				na += a[i] * a[i];
				nb += b[i] * b[i];
			}
			return d2 <= 1e-24 * std::min(na, nb);
		};
		if (approx(in.target_position, p) && approx(in.target_velocity, v)) return;
		goal_reached = false;
		for (int i = 0; i < dim; i++) {
			in.target_position[i] = p[i];
			in.target_velocity[i] = v[i];
			goal[i] = p[i];
		}
	}
	int update() {	// :118-150
		if (goal_reached) return 1;
		auto previous = out;
		const ruckig::Result r = otg.update(in, out);
		if (r == ruckig::Result::Finished) {
			double n2 = 0;
			for (int i = 0; i < dim; i++) n2 += out.new_velocity[i] * out.new_velocity[i];
			if (std::sqrt(n2) < 1e-3) {
				goal_reached = true;
			} else {
				std::vector<double> z(dim, 0.0);
				set_goal(goal.data(), z.data());
			}
			return 1;
		}
		if (r == ruckig::Result::Working) {
			out.pass_to_input(in);
			return 0;
		}
		out = previous;
		for (int i = 0; i < dim; i++) in.current_velocity[i] = in.current_acceleration[i] = 0.0;
		return (int)r;
	}
};
}  // namespace

extern "C" {
void* otg_ref_create(int dim, double dt, const double* q0) {
	auto* o = new OtgJoints(dim, dt);
	std::vector<double> z(dim, 0.0);
	o->set_goal(q0, z.data());	// reInitialize :28-41
	for (int i = 0; i < dim; i++) {
		o->out.new_position[i] = q0[i];
		o->out.new_velocity[i] = 0.0;
		o->out.new_acceleration[i] = 0.0;
	}
	o->out.pass_to_input(o->in);
	return o;
}
void otg_ref_destroy(void* p) { delete static_cast<OtgJoints*>(p); }
// jerk limit <= 0: disableJerkLimits()
void otg_ref_set_limits(void* p, double vmax, double amax, double jmax) {
	auto* o = static_cast<OtgJoints*>(p);
	for (int i = 0; i < o->dim; i++) {
		o->in.max_velocity[i] = vmax;
		o->in.max_acceleration[i] = amax;
		o->in.max_jerk[i] = jmax > 0 ? jmax : std::numeric_limits<double>::infinity();
		if (!(jmax > 0)) o->in.current_acceleration[i] = 0.0;
	}
}
void otg_ref_set_goal(void* p, const double* pos, const double* vel) { static_cast<OtgJoints*>(p)->set_goal(pos, vel); }
int otg_ref_update(void* p, double* pos, double* vel, double* acc) {
	auto* o = static_cast<OtgJoints*>(p);
	const int rc = o->update();
	for (int i = 0; i < o->dim; i++) {
		pos[i] = o->out.new_position[i];
		vel[i] = o->out.new_velocity[i];
		acc[i] = o->out.new_acceleration[i];
	}
	return rc;
}
}
