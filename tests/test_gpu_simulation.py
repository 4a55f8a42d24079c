"""the simulation side of the loop (forward dynamics + semi-implicit Euler) on the GPU against
oracle/simulation.py, and the closed loop controller <-> simulator for a batch of robots."""
import numpy as np
import pytest

from tests.osc_testlib import REL_TOL, TASK_POINTS, rng_for, sample_states

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("robot_name", ["panda", "panda_sliding_base", "rrrr", "puma_like"])
def test_integrate_matches_the_oracle(robot_name):
    import sai_primitives_b200 as sp
    from oracle import simulation as SIM
    from oracle.robots import make_chain
    from oracle.sai_model import SaiModel
    N = 24
    ch = make_chain(robot_name)
    q, dq, _ = sample_states(robot_name, N)
    tau = np.array([rng_for(i, stream=21).uniform(-5, 5, ch.n) for i in range(N)])
    robot = sp.BatchedRobot(robot_name, N)
    sim = sp.BatchedSimulation(robot, q, dq, timestep=1e-3)
    sim.setJointTorques(tau)
    sim.integrate(substeps=5)
    qg, dqg = sim.getJointPositions(), sim.getJointVelocities()
    for i in range(N):
        m = SaiModel(ch); m.setQ(q[i]); m.setDq(dq[i]); m.updateModel()
        SIM.integrate(m, tau[i], 1e-3, substeps=5)
        assert np.abs(qg[i] - m.q()).max() <= REL_TOL * max(1.0, np.abs(m.q()).max())
        assert np.abs(dqg[i] - m.dq()).max() <= REL_TOL * max(1.0, np.abs(m.dq()).max())


def test_closed_loop_batch_reaches_the_goals():
    """controller and simulator in a loop, as in examples/05-using_robot_controller: 64 Pandas with gravity compensation
    move their end-effectors to individual goals 5 cm away and hold their posture in the null space."""
    import sai_primitives_b200 as sp
    N = 64
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.12)
    dq = np.zeros_like(dq)
    link, pt = TASK_POINTS["panda"]
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)))
    jt = sp.JointTask(robot)
    ctrl = sp.RobotController(robot, [mft, jt])
    ctrl.enableGravityCompensation(True)
    x0 = mft.getCurrentPosition()
    goal = x0 + np.array([rng_for(i, stream=22).uniform(-0.03, 0.03, 3) for i in range(N)])
    mft.setGoalPosition(goal)
    sim = sp.BatchedSimulation(robot, q, dq, timestep=1e-3)
    for k in range(1500):
        tau = ctrl.step(sim.getJointPositions(), sim.getJointVelocities())
        sim.setJointTorques(tau)
        sim.integrate()
    assert (robot.status() & sp.capi.STATUS_UNHANDLED).sum() == 0
    robot.setQ(sim.getJointPositions()); robot.setDq(sim.getJointVelocities()); robot.updateModel()
    ctrl.updateControllerTaskModels(); ctrl.computeControlTorques()
    err = np.linalg.norm(mft.getCurrentPosition() - goal, axis=1)
    assert err.max() < 2e-3, err.max()
    assert np.abs(sim.getJointVelocities()).max() < 0.05


def test_urdf_registered_robot_runs_like_the_builtin_one():
    """SURVEY.md row f-3 on the GPU: a robot registered from URDF text drives the same kernels as the built-in table"""
    import sai_primitives_b200 as sp
    from oracle import robots as OR
    from tests.test_urdf_loader import _urdf_text
    N = 40
    sp.registerUrdf("urdf_panda_gpu", _urdf_text(OR.DESCRIPTIONS["panda"]()))
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.09)
    link, pt = TASK_POINTS["panda"]
    taus = []
    for name in ("panda", "urdf_panda_gpu"):
        robot = sp.BatchedRobot(name, N)
        robot.setQ(q); robot.setDq(dq); robot.updateModel()
        mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)))
        jt = sp.JointTask(robot)
        ctrl = sp.RobotController(robot, [mft, jt])
        mft.setGoalPosition(mft.getCurrentPosition() + 0.02); jt.setGoalPosition(q + 0.1)
        ctrl.updateControllerTaskModels()
        taus.append(ctrl.computeControlTorques())
        assert (robot.status() & sp.capi.STATUS_UNHANDLED).sum() == 0
    assert np.abs(taus[0] - taus[1]).max() <= 1e-11 * np.abs(taus[0]).max()
