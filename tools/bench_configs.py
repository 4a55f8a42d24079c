#!/usr/bin/env python
"""Device-resident throughput of the other BASELINE configurations (1: JointTask alone, 3: partial MotionForceTask with
closed-loop force + POPC and a JointTask), through the Python mirror of the reference API.  Not the driver's bench line:
numbers for DESIGN.md.   usage (under gpurun): python tools/bench_configs.py [robots1 robots3]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sai_primitives_b200 as sp
from sai_primitives_b200 import batched as _b
_b.JointTask.DefaultParameters.use_internal_otg = False      # BASELINE configs: internal OTG off
_b.MotionForceTask.DefaultParameters.use_internal_otg = False
import bench

def timed(ctrl, robot, q_t, dq_t, tau_t, steps=100, warm=10):
    for _ in range(warm):
        ctrl.stepDevice(q_t.data_ptr(), dq_t.data_ptr(), tau_t.data_ptr())
    robot.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        ctrl.stepDevice(q_t.data_ptr(), dq_t.data_ptr(), tau_t.data_ptr())
    robot.sync()
    return (time.perf_counter() - t0) / steps

def main():
    n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    n3 = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    dev = torch.device("cuda:0")
    link, pt = "end-effector", (0.0, 0.0, 0.07)
    # ---------------- config 1: full JointTask, kp 100 kv 20, BIE decoupling (examples/01-joint_control)
    q, dq, _, _ = bench.sample_batch(sp, n1, 0, min_ratio=0.0)
    robot = sp.BatchedRobot("panda", n1)
    jt = sp.JointTask(robot); jt.setGains(100.0, 20.0, 0.0)
    ctrl = sp.RobotController(robot, [jt])
    q_t = torch.from_numpy(np.ascontiguousarray(q.T)).to(dev); dq_t = torch.from_numpy(np.ascontiguousarray(dq.T)).to(dev)
    tau_t = torch.zeros_like(q_t)
    robot.setQ(q); robot.setDq(dq); robot.updateModel(); jt.setGoalPosition(q + 0.2)
    dt = timed(ctrl, robot, q_t, dq_t, tau_t)
    print("config 1 (JointTask alone), %d robots: %.4f ms per cycle, %.3g cycles/s" % (n1, dt * 1e3, n1 / dt))
    robot.close()
    # ---------------- config 3: ex.09: translation XYZ only, force space dim 1 about Z, closed loop + passivity, JointTask
    q, dq, _, _ = bench.sample_batch(sp, n3, 0, min_ratio=float(os.environ.get("CFG3_MIN_RATIO", "0.0")))
    robot = sp.BatchedRobot("panda", n3)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)), controlled_directions_translation=[(1, 0, 0), (0, 1, 0), (0, 0, 1)],
                             controlled_directions_rotation=[])
    jt = sp.JointTask(robot)
    ctrl = sp.RobotController(robot, [mft, jt])
    mft.parametrizeForceMotionSpaces(1, (0, 0, 1)); mft.setGoalForce(np.array([0, 0, -5.0])); mft.setClosedLoopForceControl(); mft.enablePassivity()
    mft.updateSensedForceAndMoment(np.tile([0.1, -0.2, -4.0], (n3, 1)), np.zeros((n3, 3)))
    q_t = torch.from_numpy(np.ascontiguousarray(q.T)).to(dev); dq_t = torch.from_numpy(np.ascontiguousarray(dq.T)).to(dev)
    tau_t = torch.zeros_like(q_t)
    dt = timed(ctrl, robot, q_t, dq_t, tau_t, steps=300, warm=20)
    st = robot.status()
    print("config 3 (partial MotionForceTask, closed-loop force + POPC, JointTask), %d robots, 320 cycles: %.4f ms per cycle, %.3g cycles/s;"
          " general path %.3f, unhandled %d" % (n3, dt * 1e3, n3 / dt, ((st & sp.capi.STATUS_SINGULAR_PATH) != 0).mean(), ((st & sp.capi.STATUS_UNHANDLED) != 0).sum()))

def config4(n_per_model=21845):
    """BASELINE config 4: a batch mixing Panda, PUMA-like and planar RRRR robots (one handle per model, each on its own
    stream), uniformly sampled states, i.e. with the natural share of robots inside the singular / blending band."""
    dev = torch.device("cuda:0")
    specs = {"panda": ("end-effector", (0, 0, 0.07), None, None), "puma_like": ("end-effector", (0, 0, 0.0), None, None),
             "rrrr": ("link4", (0.5, 0, 0), [(1, 0, 0), (0, 1, 0)], [(0, 0, 1)])}
    groups = []
    rng = np.random.default_rng(4)
    for name, (link, pt, dt_, dr_) in specs.items():
        desc = sp.capi.ModelDesc()
        sp.capi.load_library().osc_builtin_model(name.encode(), bench.C.byref(desc))
        n = desc.n
        lo = np.array(desc.q_lower[:n]); hi = np.array(desc.q_upper[:n])
        q = lo + (0.1 + 0.8 * rng.random((n_per_model, n))) * (hi - lo)
        dq = rng.uniform(-1, 1, (n_per_model, n))
        robot = sp.BatchedRobot(name, n_per_model)
        robot.setQ(q); robot.setDq(dq); robot.updateModel()
        mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt, dtype=float)), dt_, dr_)
        jt = sp.JointTask(robot)
        ctrl = sp.RobotController(robot, [mft, jt])
        q_t = torch.from_numpy(np.ascontiguousarray(q.T)).to(dev); dq_t = torch.from_numpy(np.ascontiguousarray(dq.T)).to(dev)
        groups.append((name, robot, ctrl, q_t, dq_t, torch.zeros_like(q_t)))
    lib = sp.capi.load_library()
    def cycle():
        for _, robot, ctrl, q_t, dq_t, tau_t in groups:
            rc = lib.osc_step_async(robot.handle, bench.C.c_void_p(q_t.data_ptr()), bench.C.c_void_p(dq_t.data_ptr()), bench.C.c_void_p(tau_t.data_ptr()),
                                    sp.capi.OSC_MEM_DEVICE)
            assert rc == 0
    for _ in range(10):
        cycle()
    for g in groups:
        g[1].sync()
    t0 = time.perf_counter()
    K = 100
    for _ in range(K):
        cycle()
    for g in groups:
        g[1].sync()
    dt = (time.perf_counter() - t0) / K
    shares = {g[0]: float(((g[1].status() & sp.capi.STATUS_SINGULAR_PATH) != 0).mean()) for g in groups}
    unh = sum(int(((g[1].status() & sp.capi.STATUS_UNHANDLED) != 0).sum()) for g in groups)
    print("config 4 (mixed batch, %d robots of each of panda / puma_like / rrrr, uniformly sampled states): %.4f ms per cycle, %.3g cycles/s;"
          " share on the general path %s, unhandled %d" % (n_per_model, dt * 1e3, 3 * n_per_model / dt, shares, unh))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "config4":
        config4(int(sys.argv[2]) if len(sys.argv) > 2 else 21845)
        sys.exit(0)
    main()
