// TEST INFRASTRUCTURE (oracle/_ref): C entry points around the REFERENCE's own POPCExplicitForceControl class
// (/root/reference/src/helper_modules/POPCExplicitForceControl.cpp, compiled where it lies by oracle/Makefile against
// oracle/eigen_standin).  Used only to generate tests/golden/popc_reference.npz and to check the numpy restatement.
#include "POPCExplicitForceControl.h"

extern "C" {
void* popc_ref_create(double loop_timestep) { return new SaiPrimitives::POPCExplicitForceControl(loop_timestep); }
void popc_ref_destroy(void* p) { delete static_cast<SaiPrimitives::POPCExplicitForceControl*>(p); }
void popc_ref_enable(void* p, int on) {
	auto* c = static_cast<SaiPrimitives::POPCExplicitForceControl*>(p);
	if (on) c->enable(); else c->disable();
}
void popc_ref_reinitialize(void* p) { static_cast<SaiPrimitives::POPCExplicitForceControl*>(p)->reInitialize(); }
// kv: 3 x 3 row major
void popc_ref_step(void* p, const double* fd, const double* fs, const double* vcl, const double* vr, const double* kv, double kff, double* out) {
	auto* c = static_cast<SaiPrimitives::POPCExplicitForceControl*>(p);
	Eigen::Matrix3d K;
	for (int r = 0; r < 3; r++)
		for (int q = 0; q < 3; q++) K(r, q) = kv[3 * r + q];
	const Eigen::Vector3d o = c->computePassivitySaturatedForce(Eigen::Vector3d(fd[0], fd[1], fd[2]), Eigen::Vector3d(fs[0], fs[1], fs[2]),
																 Eigen::Vector3d(vcl[0], vcl[1], vcl[2]), Eigen::Vector3d(vr[0], vr[1], vr[2]), K, kff);
	out[0] = o(0); out[1] = o(1); out[2] = o(2);
}
}
