// Host-side decisions of the launcher (csrc/osc_launch.h) for tests/test_launch_logic.py: which hierarchies the specialised
// fused kernel (and with it the FP32 mode) covers, and when a cycle takes the split blending path.
#include "../../sai_primitives_b200/csrc/osc_launch.h"
#include <cstring>

static void base_program(OscProgram& P) {
	std::memset(&P, 0, sizeof(P));
	P.model.n = 7;
	for (int j = 0; j < 7; j++) {
		P.model.axis[j][2] = 1.0;
		P.model.inertia[j][0] = P.model.inertia[j][3] = P.model.inertia[j][5] = 0.3;	// principal moments well above the BIE threshold
	}
	P.n_robots = 65536;
	P.n_tasks = 2;
	P.tasks[0].type = OSC_TASK_MOTION_FORCE;
	P.tasks[1].type = OSC_TASK_JOINT;
	P.mft[0].body = 6;
	P.mft[0].full = 1;
	P.mft[0].rank = 6;
	P.mft[0].p.dynamic_decoupling_type = OSC_BOUNDED_INERTIA_ESTIMATES;
	P.mft[0].p.bie_threshold = 0.1;
	P.jt[0].full = 1;
	P.jt[0].p.dynamic_decoupling_type = OSC_BOUNDED_INERTIA_ESTIMATES;
	P.jt[0].p.bie_threshold = 0.1;
}

// bit 0: eligible for the specialised kernel, bit 1: pure motion control
extern "C" int spec_case(int which) {
	OscProgram P;
	base_program(P);
	switch (which) {
	case 0: break;															 // the flagship configuration
	case 1: P.mft[0].p.force_space_dimension = 1; break;					 // force control: specialised, not "motion"
	case 2: P.model.axis[3][2] = 0.0; P.model.axis[3][0] = 1.0; break;		 // an off-axis joint
	case 3: P.model.jtype[0] = 1; break;									 // a prismatic joint
	case 4: P.jt[0].p.use_velocity_saturation = 1; break;					 // velocity saturation of the joint task
	case 5: P.mft[0].p.bie_threshold = 0.7; break;							 // two diagonal entries of M may fall below the threshold (0.3, 0.6)
	case 6: P.mft[0].p.use_velocity_saturation = 1; break;					 // of the motion-force task: specialised, general control law
	case 7: P.n_robots = 40000000; break;									 // element indices beyond 32 bits
	case 8: P.mft[0].p.closed_loop_force_control = 1; break;
	}
	bool motion = false;
	const bool ok = osc::cycle_spec_eligible(P, true, &motion);
	return (ok ? 1 : 0) | ((ok && motion) ? 2 : 0);
}

extern "C" int split_case(int which) {
	OscProgram P;
	base_program(P);
	static double scratch;
	static uint32_t done[4];
	P.blend_split_on = 1;
	P.blend_scratch = &scratch;
	P.general_done = done;
	switch (which) {
	case 0: break;
	case 1: P.blend_split_on = 0; break;		// the host hint does not say "many"
	case 2: P.blend_scratch = nullptr; break;	// nothing allocated
	case 3: P.general_done = nullptr; break;	// handle without cross-cycle pipelining
	case 4: P.general_grid_small = 1; break;	// the hint says "clean"
	case 5: P.precision_fp32 = 1; break;		// the FP32 kernel has no parking instantiation
	case 6: P.mft[0].full = 0; break;			// partial task
	case 7: P.mft[0].rank = 3; break;
	case 8: P.tasks[0].type = OSC_TASK_JOINT; break;
	}
	return osc::blend_split_selected(P) ? 1 : 0;
}

extern "C" double min_eig(const double* I6) { return osc::min_eig_sym3(I6); }
