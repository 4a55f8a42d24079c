"""The optional single-precision mode of the north star ("an optional FP32 mode is held to 1e-4 relative and reported
separately"): the fused kernel of config 2 (Panda, full six-dof MotionForceTask + JointTask in its null space through
RobotController) computed in FP32 on the FP64 state, against the reference's compiled control law (double precision,
RobotController.cpp:75-120) and against the FP64 mode of the same library.  Tolerance: 1e-4 relative per robot."""
import numpy as np
import pytest

from tests.osc_testlib import TASK_POINTS, OracleBatch, rel_err, rng_for, rot_exp, sample_states

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4


def _build(sp, q, dq, dec, with_jt=True, gravity=False):
    N = q.shape[0]
    link, pt = TASK_POINTS["panda"]
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)))
    mft.setDynamicDecouplingType(dec)
    mft.setPosControlGains(100.0, 20.0, 2.0); mft.setOriControlGains(200.0, 28.3, 1.0)
    tasks = [mft]
    jt = None
    if with_jt:
        jt = sp.JointTask(robot); jt.setDynamicDecouplingType(dec); jt.setGains(50.0, 14.0, 1.0)
        tasks.append(jt)
    ctrl = sp.RobotController(robot, tasks)
    if gravity:
        ctrl.enableGravityCompensation(True)
    return robot, mft, jt, ctrl


def _goals(mft, jt, q, N):
    x0 = mft.getCurrentPosition(); R0 = mft.getCurrentOrientation()
    xd = np.zeros((N, 3)); Rd = np.zeros((N, 3, 3)); vd = np.zeros((N, 3)); gq = np.zeros((N, 7))
    for i in range(N):
        g = rng_for(i, stream=9)
        xd[i] = x0[i] + g.uniform(-0.05, 0.05, 3); Rd[i] = R0[i] @ rot_exp(g.uniform(-0.2, 0.2, 3)); vd[i] = g.uniform(-0.1, 0.1, 3)
        gq[i] = q[i] + g.uniform(-0.2, 0.2, 7)
    mft.setGoalPosition(xd); mft.setGoalOrientation(Rd); mft.setGoalLinearVelocity(vd)
    if jt is not None:
        jt.setGoalPosition(gq)
    return xd, Rd, vd, gq


@pytest.mark.parametrize("dec", [0, 1, 2])
@pytest.mark.parametrize("with_jt,gravity", [(True, False), (True, True), (False, False)])
def test_fp32_mode_config2_against_the_reference(dec, with_jt, gravity):
    import sai_primitives_b200 as sp
    N, K = 256, 6
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.1)
    robot, mft, jt, ctrl = _build(sp, q, dq, dec, with_jt, gravity)
    ctrl.setPrecision("fp32")
    assert ctrl.getPrecision() == "fp32"
    xd, Rd, vd, gq = _goals(mft, jt, q, N)
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    link, pt = TASK_POINTS["panda"]
    omft = ob.add_mft(link, (np.eye(3), np.array(pt))); ojt = ob.add_jt() if with_jt else None
    ob.finalize()
    if gravity:
        for c in ob.controllers:
            c.enableGravityCompensation(True)
    for i in range(N):
        omft[i].setDynamicDecouplingType(dec); omft[i].setPosControlGains(100.0, 20.0, 2.0); omft[i].setOriControlGains(200.0, 28.3, 1.0)
        omft[i].setGoalPosition(xd[i]); omft[i].setGoalOrientation(Rd[i]); omft[i].setGoalLinearVelocity(vd[i])
        if with_jt:
            ojt[i].setDynamicDecouplingType(dec); ojt[i].setGains(50.0, 14.0, 1.0); ojt[i].setGoalPosition(gq[i])
    worst = 0.0
    for k in range(K):      # integral gains: every cycle reads what the previous one accumulated (in FP32, stored as FP64)
        ctrl.updateControllerTaskModels()
        tau = ctrl.computeControlTorques()
        ref = ob.cycle()
        st = robot.status()
        assert (st & sp.capi.STATUS_UNHANDLED).sum() == 0
        err = rel_err(tau, ref)
        worst = max(worst, float(err.max()))
        assert err.max() < FP32_TOL, (k, float(err.max()))
        qk = q + 0.002 * (k + 1) * dq
        robot.setQ(qk); robot.updateModel(); ob.set_state(qk, dq)
    assert worst > 1e-9          # it really was single precision


def test_fp32_mode_against_fp64_mode_and_handovers_stay_fp64():
    """unfiltered states: the robots inside the singularity band leave the FP32 kernel for the FP64 general path and keep the
    1e-9 agreement with the FP64 mode; the others agree to 1e-4"""
    import sai_primitives_b200 as sp
    N = 2048
    q, dq, _ = sample_states("panda", N)
    out = {}
    for mode in ("fp64", "fp32"):
        robot, mft, jt, ctrl = _build(sp, q, dq, 1)
        ctrl.setPrecision(mode)
        _goals(mft, jt, q, N)
        for k in range(3):
            ctrl.updateControllerTaskModels()
            tau = ctrl.computeControlTorques()
        out[mode] = (tau, robot.status())
        robot.close()
    st64, st32 = out["fp64"][1], out["fp32"][1]
    sing64 = (st64 & sp.capi.STATUS_SINGULAR_PATH) != 0
    sing32 = (st32 & sp.capi.STATUS_SINGULAR_PATH) != 0
    assert sing64.mean() > 0.3 and (sing64 != sing32).mean() < 0.01     # the branch decision may differ in a thin band only
    both = sing64 & sing32
    err = rel_err(out["fp32"][0], out["fp64"][0])
    assert err[both].max() < 1e-9
    neither = ~sing64 & ~sing32
    assert err[neither].max() < FP32_TOL and err[neither].max() > 1e-9


def test_fp32_mode_refuses_other_hierarchies():
    import sai_primitives_b200 as sp
    N = 16
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.1)
    link, pt = TASK_POINTS["panda"]
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)), [(1, 0, 0), (0, 1, 0), (0, 0, 1)], [])     # partial task
    jt = sp.JointTask(robot)
    ctrl = sp.RobotController(robot, [mft, jt])
    ctrl.setPrecision("fp32")
    ctrl.updateControllerTaskModels()
    with pytest.raises(NotImplementedError):        # OSC_ERR_UNSUPPORTED: no silent fallback to FP64
        ctrl.computeControlTorques()
    ctrl.setPrecision("fp64")
    assert np.isfinite(ctrl.computeControlTorques()).all()
    robot.close()
    # a dof without a single-precision kernel is refused when the mode is requested
    robot = sp.BatchedRobot("puma_like", 4)
    mft = sp.MotionForceTask(robot, "end-effector", (np.eye(3), np.zeros(3)))
    ctrl = sp.RobotController(robot, [mft])
    with pytest.raises(NotImplementedError):
        ctrl.setPrecision("fp32")
