"""Golden fixtures (tests/golden/*.npz, frozen oracle outputs -- see tests/golden/generate.py; the
reference itself ships none).  CPU: the C++ restatement reproduces them.  GPU: the CUDA path
reproduces them through the C ABI without touching oracle/ at run time."""
import os

import numpy as np
import pytest

from tests.osc_testlib import REL_TOL, TASK_POINTS, rel_err

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


# ------------------------------------------------------------------ CPU: C++ oracle vs goldens
def test_cpp_oracle_reproduces_config1_and_config2():
    from oracle.cpp_ref import CppOracleBatch
    d = load("config1_joint_task.npz")
    N = d["q"].shape[0]
    cb = CppOracleBatch("panda", N); cb.set_state(d["q"], d["dq"])
    tj = cb.add_jt(); cb.jt_set_gains(tj, 100.0, 20.0, 3.0); cb.jt_set_goals(tj, d["qd"])
    for k in range(3):
        assert rel_err(cb.cycle(), d["tau"][k]).max() < 1e-9
    d = load("config2_osc_nullspace.npz")
    N = d["q"].shape[0]
    link, pt = TASK_POINTS["panda"]
    cb = CppOracleBatch("panda", N); cb.set_state(d["q"], d["dq"])
    tm = cb.add_mft(link, (np.eye(3), np.array(pt))); tj = cb.add_jt()
    cb.mft_set_goals(tm, d["xd"], d["Rd"], d["vd"], d["wd"], d["ad"], d["ald"]); cb.jt_set_goals(tj, d["qd"])
    assert d["singular"].sum() >= 8      # the fixture covers the blending branch
    for k in range(3):
        assert rel_err(cb.cycle(), d["tau"][k]).max() < 1e-9


def test_cpp_oracle_reproduces_force_popc_trajectory():
    from oracle.cpp_ref import CppOracleBatch
    d = load("config3_force_popc.npz")
    N = d["q"].shape[0]
    link, pt = TASK_POINTS["panda"]
    cb = CppOracleBatch("panda", N); cb.set_state(d["q"], d["dq"])
    tm = cb.add_mft(link, (np.eye(3), np.array(pt)), [(1, 0, 0), (0, 1, 0), (0, 0, 1)], []); tj = cb.add_jt()
    assert rel_err(cb.cycle(), d["tau0"]).max() < 1e-9
    cb.mft_force_setup(tm, fdim=1, faxis=(0, 0, 1), cl_force=True, passivity=True)
    cb.mft_set_force_goals(tm, np.tile([0, 0, -5.0], (N, 1)), np.zeros((N, 3)))
    for k in range(d["tau"].shape[0]):
        cb.mft_update_sensed(tm, d["sensed_force"][k], d["sensed_moment"][k])
        assert rel_err(cb.cycle(), d["tau"][k]).max() < 1e-9, k
    assert d["Rc"].min() < 1.0     # the passivity controller really engaged in the fixture


def test_cpp_oracle_reproduces_mixed_dof():
    from oracle.cpp_ref import CppOracleBatch
    d = load("config4_mixed_dof.npz")
    for name, dt_, dr_ in (("rrrr", [(1, 0, 0), (0, 1, 0)], [(0, 0, 1)]), ("puma_like", None, None)):
        q = d[name + "_q"]; N = q.shape[0]
        link, pt = TASK_POINTS[name]
        cb = CppOracleBatch(name, N); cb.set_state(q, d[name + "_dq"])
        tm = cb.add_mft(link, (np.eye(3), np.array(pt)), dt_, dr_); tj = cb.add_jt()
        cb.mft_set_goals(tm, *[d[name + "_" + k] for k in ("xd", "Rd", "vd", "wd", "ad", "ald")]); cb.jt_set_goals(tj, d[name + "_qd"])
        for k in range(3):
            assert rel_err(cb.cycle(), d[name + "_tau"][k]).max() < 1e-9


# ------------------------------------------------------------------ GPU: CUDA path vs goldens
def _gpu_osc(sp, name, q, dq, dt_, dr_, goals):
    N = q.shape[0]
    link, pt = TASK_POINTS[name]
    robot = sp.BatchedRobot(name, N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)), dt_, dr_)
    jt = sp.JointTask(robot)
    ctrl = sp.RobotController(robot, [mft, jt])
    if goals is not None:
        mft.setGoalPosition(goals["xd"]); mft.setGoalOrientation(goals["Rd"]); mft.setGoalLinearVelocity(goals["vd"])
        mft.setGoalAngularVelocity(goals["wd"]); mft.setGoalLinearAcceleration(goals["ad"]); mft.setGoalAngularAcceleration(goals["ald"])
        jt.setGoalPosition(goals["qd"])
    return robot, mft, jt, ctrl


@pytest.mark.gpu
def test_gpu_reproduces_config1():
    import sai_primitives_b200 as sp
    d = load("config1_joint_task.npz")
    N = d["q"].shape[0]
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(d["q"]); robot.setDq(d["dq"]); robot.updateModel()
    jt = sp.JointTask(robot); jt.setGains(100.0, 20.0, 3.0); jt.setGoalPosition(d["qd"])
    ctrl = sp.RobotController(robot, [jt])
    for k in range(3):
        ctrl.updateControllerTaskModels()
        assert rel_err(ctrl.computeControlTorques(), d["tau"][k]).max() < REL_TOL


@pytest.mark.gpu
def test_gpu_reproduces_config2_and_mixed_dof():
    import sai_primitives_b200 as sp
    d = load("config2_osc_nullspace.npz")
    robot, mft, jt, ctrl = _gpu_osc(sp, "panda", d["q"], d["dq"], None, None, d)
    assert np.abs(mft.getCurrentPosition() - d["x0"]).max() < 1e-13
    for k in range(3):
        ctrl.updateControllerTaskModels()
        tau = ctrl.computeControlTorques()
        st = robot.status()
        assert (st & sp.capi.STATUS_UNHANDLED).sum() == 0
        on_svd_path = (st & sp.capi.STATUS_SINGULAR_PATH) != 0
        assert on_svd_path[d["singular"]].all()               # never treats a singular robot as non-singular
        assert on_svd_path[~d["singular"]].mean() < 0.2       # the sound test rejects only a thin band
        assert rel_err(tau, d["tau"][k]).max() < REL_TOL      # blending branch, type-1/type-2 strategies included
    m = load("config4_mixed_dof.npz")
    for name, dt_, dr_ in (("rrrr", [(1, 0, 0), (0, 1, 0)], [(0, 0, 1)]), ("puma_like", None, None)):
        goals = {k: m[name + "_" + k] for k in ("xd", "Rd", "vd", "wd", "ad", "ald", "qd")}
        robot, mft, jt, ctrl = _gpu_osc(sp, name, m[name + "_q"], m[name + "_dq"], dt_, dr_, goals)
        for k in range(3):
            ctrl.updateControllerTaskModels()
            tau = ctrl.computeControlTorques()
            st = robot.status()
            assert (st & sp.capi.STATUS_UNHANDLED).sum() == 0
            assert ((st & sp.capi.STATUS_SINGULAR_PATH) != 0)[m[name + "_singular"]].all()
            assert rel_err(tau, m[name + "_tau"][k]).max() < REL_TOL


@pytest.mark.gpu
def test_gpu_reproduces_force_popc_trajectory():
    """BASELINE config 3: closed-loop force with POPC passivity state carried on the device for 320 cycles."""
    import sai_primitives_b200 as sp
    d = load("config3_force_popc.npz")
    N = d["q"].shape[0]
    robot, mft, jt, ctrl = _gpu_osc(sp, "panda", d["q"], d["dq"], [(1, 0, 0), (0, 1, 0), (0, 0, 1)], [], None)
    ctrl.updateControllerTaskModels()
    assert rel_err(ctrl.computeControlTorques(), d["tau0"]).max() < REL_TOL
    assert mft.parametrizeForceMotionSpaces(1, (0, 0, 1)) is True
    mft.setGoalForce(np.array([0, 0, -5.0])); mft.setClosedLoopForceControl(); mft.enablePassivity()
    worst = 0.0
    for k in range(d["tau"].shape[0]):
        mft.updateSensedForceAndMoment(d["sensed_force"][k], d["sensed_moment"][k])
        ctrl.updateControllerTaskModels()
        tau = ctrl.computeControlTorques()
        worst = max(worst, rel_err(tau, d["tau"][k]).max())
        assert worst < REL_TOL, k
    popc = mft._get(sp.capi.MFT_POPC_STATE)
    assert np.abs(popc[:, 2] - d["Rc"][-1]).max() < 1e-9
    assert (robot.status() & sp.capi.STATUS_POPC_OVERFLOW).sum() == 0
