// The reference's example 05 control loop (examples/05-using_robot_controller/05-using_robot_controller.cpp:103-196),
// written against the batched C++ mirror: MotionForceTask at end-effector (0,0,0.07) + JointTask in its null space,
// RobotController, for N robots.  Usage: host_mirror_demo <state.bin> <N> <tau_out.bin>
//   state.bin: q (7 x N, SoA) then dq (7 x N, SoA), float64.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "sai_b200/sai_primitives_batched.hpp"

using namespace SaiPrimitivesB200;

int main(int argc, char** argv) {
	if (argc < 4) return 2;
	const int64_t N = std::atoll(argv[2]);
	const int dof = 7;
	std::vector<double> q((size_t)dof * N), dq((size_t)dof * N);
	FILE* f = std::fopen(argv[1], "rb");
	if (!f || std::fread(q.data(), 8, q.size(), f) != q.size() || std::fread(dq.data(), 8, dq.size(), f) != dq.size()) return 3;
	std::fclose(f);
	try {
		auto robot = std::make_shared<BatchedRobot>("panda", N);
		robot->setState(q.data(), dq.data());
		auto motion_force_task = std::make_shared<MotionForceTask>(robot, "end-effector", Affine::Translation(0.0, 0.0, 0.07));
		motion_force_task->disableInternalOtg();
		auto joint_task = std::make_shared<JointTask>(robot);
		joint_task->disableInternalOtg();  // like the reference's examples (examples/01-joint_control/01-joint_control.cpp:136)
		std::vector<std::shared_ptr<TemplateTask>> task_list = {motion_force_task, joint_task};
		auto robot_controller = std::make_unique<RobotController>(robot, task_list);

		// the reference throws std::invalid_argument for negative gains (MotionForceTask.cpp:583-587)
		bool threw = false;
		try {
			motion_force_task->setPosControlGains(-1.0, 20.0);
		} catch (const std::invalid_argument&) {
			threw = true;
		}
		if (!threw) return 4;
		// and a task after a full joint task is rejected (RobotController.cpp:45-51) -- here: a second controller on a finalized robot
		motion_force_task->setGoalLinearVelocity(Vec3{0.01, -0.02, 0.03});
		std::vector<double> goal(7, 0.1);
		joint_task->setGoalVelocity(goal);

		robot_controller->updateControllerTaskModels();
		std::vector<double> tau = robot_controller->computeControlTorques();
		FILE* o = std::fopen(argv[3], "wb");
		std::fwrite(tau.data(), 8, tau.size(), o);
		std::fclose(o);
		size_t unhandled = 0, singular = 0;
		for (uint32_t s : robot->status()) {
			unhandled += (s & OSC_STATUS_UNHANDLED) ? 1 : 0;
			singular += (s & OSC_STATUS_SINGULAR_PATH) ? 1 : 0;
		}
		// TemplateTask::getTaskAndPreviousNullspace of the last task of a full hierarchy on a 7-dof arm: the zero matrix
		std::vector<double> Nall((size_t)49 * N);
		joint_task->getTaskAndPreviousNullspace(Nall.data());
		double nmax = 0.0;
		for (double v : Nall) nmax = std::max(nmax, std::fabs(v));
		std::printf("ok %lld robots, %zu unhandled, %zu on the singular path, |N_joint N_prec| max %.1e\n", (long long)N, unhandled, singular, nmax);
	} catch (const std::exception& e) {
		std::fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	return 0;
}
