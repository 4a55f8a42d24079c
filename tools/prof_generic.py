import sys, numpy as np, torch
sys.path.insert(0, ".")
import sai_primitives_b200 as sp, bench
name, link, pt = "puma_like", "end-effector", (0.0, 0.0, 0.0)
n_rob = 8192
desc = sp.capi.ModelDesc(); sp.capi.load_library().osc_builtin_model(name.encode(), bench.C.byref(desc))
n = desc.n; lo = np.array(desc.q_lower[:n]); hi = np.array(desc.q_upper[:n])
rng = np.random.default_rng(4)
q = lo + (0.1 + 0.8 * rng.random((n_rob, n))) * (hi - lo); dq = rng.uniform(-1, 1, (n_rob, n))
robot = sp.BatchedRobot(name, n_rob); robot.setQ(q); robot.setDq(dq); robot.updateModel()
mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt))); jt = sp.JointTask(robot)
ctrl = sp.RobotController(robot, [mft, jt])
for _ in range(3):
    ctrl.updateControllerTaskModels(); ctrl.computeControlTorques()
print("general path share", ((robot.status() & sp.capi.STATUS_SINGULAR_PATH) != 0).mean())
