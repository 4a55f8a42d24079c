// Host build of the product's OTG core (sai_primitives_b200/csrc/osc_otg.h, the same header the CUDA kernels include) with C
// entry points, so that tests/test_otg_core.py can check it on the CPU against the reference's own Ruckig + OTG_joints.cpp.
#include <new>

#include "../../sai_primitives_b200/csrc/osc_otg.h"

namespace {
struct Probe {
	int k;
	double dt;
	double vmax[8], amax[8];
	otg::JointsOtg<8> o;
	otg::Calculator<8> calc;
};
}  // namespace

extern "C" {
void* otgp_create(int k, double dt, const double* q0) {
	Probe* p = new Probe();
	p->k = k;
	p->dt = dt;
	p->o.flags = 0;
	p->o.time = p->o.duration = 0.0;
	for (int i = 0; i < 8; i++) p->o.target_pos[i] = p->o.target_vel[i] = 0.0;
	p->o.reinitialize(k, q0);
	return p;
}
void otgp_destroy(void* h) { delete static_cast<Probe*>(h); }
void otgp_set_limits(void* h, const double* vmax, const double* amax) {
	Probe* p = static_cast<Probe*>(h);
	for (int i = 0; i < p->k; i++) {
		p->vmax[i] = vmax[i];
		p->amax[i] = amax[i];
		p->o.in_acc[i] = 0.0;  // OTG_joints::disableJerkLimits (:89-92)
	}
	p->o.flags |= otg::OTG_DIRTY;
}
void otgp_set_goal(void* h, const double* pos, const double* vel) {
	Probe* p = static_cast<Probe*>(h);
	p->o.set_goal(p->k, pos, vel);
}
void otgp_reinitialize(void* h, const double* pos) {
	Probe* p = static_cast<Probe*>(h);
	p->o.reinitialize(p->k, pos);
}
int otgp_update(void* h, double* pos, double* vel, double* acc) {
	Probe* p = static_cast<Probe*>(h);
	p->o.update(p->k, p->dt, p->vmax, p->amax, &p->calc);
	for (int i = 0; i < p->k; i++) {
		pos[i] = p->o.out_pos[i];
		vel[i] = p->o.out_vel[i];
		acc[i] = p->o.out_acc[i];
	}
	return p->o.flags;
}
}

// ---- OTG_6dof_cartesian
namespace {
struct CartProbe {
	double dt;
	double vmax[6], amax[6];
	otg::CartesianOtg o;
	otg::Calculator<6> calc;
};
}  // namespace
extern "C" {
void* otgc_create(double dt, const double* pos, const double* R, double vlin, double alin, double vang, double aang) {
	CartProbe* p = new CartProbe();
	p->dt = dt;
	p->o.construct(pos, R);
	for (int i = 0; i < 3; i++) {
		p->vmax[i] = vlin;
		p->amax[i] = alin;
		p->vmax[3 + i] = vang;
		p->amax[3 + i] = aang;
	}
	return p;
}
void otgc_destroy(void* h) { delete static_cast<CartProbe*>(h); }
void otgc_set_goal(void* h, const double* pos, const double* v, const double* R, const double* w) {
	CartProbe* p = static_cast<CartProbe*>(h);
	p->o.set_goal_linear(pos, v);
	p->o.set_goal_angular(R, w);
}
void otgc_reinitialize(void* h, const double* pos, const double* R) { static_cast<CartProbe*>(h)->o.reinitialize(pos, R); }
// out: pos 3, R 9, v 3, w 3, a 3, alpha 3
int otgc_update(void* h, double* out) {
	CartProbe* p = static_cast<CartProbe*>(h);
	p->o.update(p->dt, p->vmax, p->amax, &p->calc);
	p->o.desired(out, out + 3, out + 12, out + 15, out + 18, out + 21);
	return p->o.g.flags;
}
}
