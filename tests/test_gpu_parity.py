"""Parity of the fused CUDA control cycle with the CPU restatement of the reference
(oracle/primitives.py), through the C ABI.  Tolerance: 1e-9 relative per robot
(BASELINE.json north_star), FP64."""
import numpy as np
import pytest

from tests.osc_testlib import REL_TOL, TASK_POINTS, OracleBatch, rel_err, rng_for, rot_exp, sample_states

pytestmark = pytest.mark.gpu

DEC = [0, 1, 2]  # FULL_DYNAMIC_DECOUPLING, BOUNDED_INERTIA_ESTIMATES, IMPEDANCE


def _set_joint_goals(jt_gpu, jt_or, q, n, with_vel=False):
    N = q.shape[0]
    gp = np.zeros((N, n)); gv = np.zeros((N, n)); ga = np.zeros((N, n))
    for i in range(N):
        g = rng_for(i, stream=1)
        gp[i] = q[i] + g.uniform(-0.2, 0.2, n)
        if with_vel:
            gv[i] = g.uniform(-0.1, 0.1, n); ga[i] = g.uniform(-0.5, 0.5, n)
        jt_or[i].setGoalPosition(gp[i]); jt_or[i].setGoalVelocity(gv[i]); jt_or[i].setGoalAcceleration(ga[i])
    jt_gpu.setGoalPosition(gp); jt_gpu.setGoalVelocity(gv); jt_gpu.setGoalAcceleration(ga)


def _set_mft_goals(mft_gpu, mft_or, N):
    x0 = mft_gpu.getCurrentPosition(); R0 = mft_gpu.getCurrentOrientation()
    xd = np.zeros((N, 3)); Rd = np.zeros((N, 3, 3)); vd = np.zeros((N, 3)); wd = np.zeros((N, 3)); ad = np.zeros((N, 3)); ald = np.zeros((N, 3))
    for i in range(N):
        g = rng_for(i, stream=2)
        xd[i] = x0[i] + g.uniform(-0.05, 0.05, 3)
        Rd[i] = R0[i] @ rot_exp(g.uniform(-0.2, 0.2, 3))
        vd[i] = g.uniform(-0.1, 0.1, 3); wd[i] = g.uniform(-0.1, 0.1, 3)
        ad[i] = g.uniform(-0.5, 0.5, 3); ald[i] = g.uniform(-0.5, 0.5, 3)
        t = mft_or[i]
        # goals are built from the GPU's own current pose: check the oracle agrees on it first
        assert np.abs(t._current_position - x0[i]).max() < 1e-12
        assert np.abs(t._current_orientation - R0[i]).max() < 1e-12
        t.setGoalPosition(xd[i]); t.setGoalOrientation(Rd[i]); t.setGoalLinearVelocity(vd[i])
        t.setGoalAngularVelocity(wd[i]); t.setGoalLinearAcceleration(ad[i]); t.setGoalAngularAcceleration(ald[i])
    mft_gpu.setGoalPosition(xd); mft_gpu.setGoalOrientation(Rd); mft_gpu.setGoalLinearVelocity(vd)
    mft_gpu.setGoalAngularVelocity(wd); mft_gpu.setGoalLinearAcceleration(ad); mft_gpu.setGoalAngularAcceleration(ald)


@pytest.mark.parametrize("dec", DEC)
@pytest.mark.parametrize("robot_name", ["panda", "rrrr"])
def test_config1_joint_task_alone(robot_name, dec):
    """BASELINE config 1: single full JointTask, kp 100 kv 20 (examples/01-joint_control/...cpp:133)."""
    import sai_primitives_b200 as sp
    N = 96
    q, dq, _ = sample_states(robot_name, N)
    n = q.shape[1]
    robot = sp.BatchedRobot(robot_name, N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    jt = sp.JointTask(robot)
    jt.setGains(100.0, 20.0, 3.0); jt.setDynamicDecouplingType(dec)
    ctrl = sp.RobotController(robot, [jt])
    ob = OracleBatch(robot_name, N); ob.set_state(q, dq)
    ojt = ob.add_jt()
    for t in ojt:
        t.setGains(100.0, 20.0, 3.0); t.setDynamicDecouplingType(dec)
    ob.finalize()
    _set_joint_goals(jt, ojt, q, n, with_vel=True)
    for cycle in range(3):   # the integrator state must evolve identically
        ctrl.updateControllerTaskModels()
        tau = ctrl.computeControlTorques()
        ref = ob.cycle()
        assert rel_err(tau, ref).max() < REL_TOL
    assert (robot.status() == 0).all()


@pytest.mark.parametrize("use_prev", [True, False])
@pytest.mark.parametrize("dec", DEC)
def test_config2_panda_osc_with_nullspace_joint_task(dec, use_prev):
    """BASELINE config 2: MotionForceTask 6-DoF at end-effector (0,0,0.07) + JointTask in its null
    space through RobotController (examples/05) and the manual sum of examples/04."""
    import sai_primitives_b200 as sp
    N = 128
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.075)
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    link, pt = TASK_POINTS["panda"]
    mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)))
    jt = sp.JointTask(robot)
    mft.setDynamicDecouplingType(dec); jt.setDynamicDecouplingType(dec)
    ctrl = sp.RobotController(robot, [mft, jt], use_previous_torques=use_prev)
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    omft = ob.add_mft(link, (np.eye(3), np.array(pt))); ojt = ob.add_jt()
    for a, b in zip(omft, ojt):
        a.setDynamicDecouplingType(dec); b.setDynamicDecouplingType(dec)
    ob.finalize()
    _set_mft_goals(mft, omft, N); _set_joint_goals(jt, ojt, q, 7)
    ctrl.updateControllerTaskModels()
    tau = ctrl.computeControlTorques()
    ref = ob.cycle(use_prev=use_prev)
    st = robot.status()
    assert (st & sp.capi.STATUS_UNHANDLED).sum() == 0
    fast = (st & sp.capi.STATUS_SINGULAR_PATH) == 0
    # the sound test may send a thin band of non-singular robots to the SVD path; nearly all stay on the fast path
    assert fast.mean() > 0.9
    for i in np.nonzero(fast)[0]:
        assert len(omft[i]._singularity_handler._singularity_types) == 0
    assert rel_err(tau, ref).max() < REL_TOL


@pytest.mark.parametrize("N", [1, 31, 333])
def test_config2_ragged_batch_sizes_multi_cycle(N):
    """batch sizes that are not a multiple of the warp / block size (and odd, so that the component rows of the SoA state
    are only 8-byte aligned), several cycles with integral gains so that the integrators the specialised kernel stages
    through shared memory carry over"""
    import sai_primitives_b200 as sp
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.09)
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    link, pt = TASK_POINTS["panda"]
    mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)))
    jt = sp.JointTask(robot)
    mft.setPosControlGains(100.0, 20.0, 35.0); mft.setOriControlGains(200.0, 28.3, 25.0); jt.setGains(50.0, 14.0, 12.0)
    ctrl = sp.RobotController(robot, [mft, jt])
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    omft = ob.add_mft(link, (np.eye(3), np.array(pt))); ojt = ob.add_jt()
    for a, b in zip(omft, ojt):
        a.setPosControlGains(100.0, 20.0, 35.0); a.setOriControlGains(200.0, 28.3, 25.0); b.setGains(50.0, 14.0, 12.0)
    ob.finalize()
    _set_mft_goals(mft, omft, N); _set_joint_goals(jt, ojt, q, 7)
    for cycle in range(4):
        ctrl.updateControllerTaskModels()
        tau = ctrl.computeControlTorques()
        ref = ob.cycle()
        assert (robot.status() & sp.capi.STATUS_UNHANDLED).sum() == 0
        assert rel_err(tau, ref).max() < REL_TOL, cycle
        q = q + 0.001 * dq
        robot.setQ(q); robot.updateModel(); ob.set_state(q, dq)


def test_config2_task_on_an_inner_link():
    """the motion-force task on link6 of the Panda: joint 7 does not move the task frame, its Jacobian column is zero
    (the masked iteration of the rolled Jacobian pass in the specialised kernel)"""
    import sai_primitives_b200 as sp
    from oracle.robots import make_chain
    from oracle.sai_model import SaiModel
    N = 160
    link, pt = "link6", (0.05, 0.0, 0.02)
    ch = make_chain("panda"); model = SaiModel(ch)
    q = np.zeros((N, 7)); dq = np.zeros((N, 7))
    for i in range(N):                                   # rejection sampling on this task's own Jacobian
        g = rng_for(i, stream=31)
        for _ in range(2000):
            qi = ch.q_lower + (0.1 + 0.8 * g.random(7)) * (ch.q_upper - ch.q_lower)
            model.setQ(qi); model.updateKinematics()
            sv = np.linalg.svd(model.J(link, pt), compute_uv=False)
            if sv[5] / sv[0] >= 0.09:
                break
        q[i] = qi; dq[i] = g.uniform(-1, 1, 7)
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)))
    jt = sp.JointTask(robot)
    ctrl = sp.RobotController(robot, [mft, jt])
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    omft = ob.add_mft(link, (np.eye(3), np.array(pt))); ojt = ob.add_jt(); ob.finalize()
    _set_mft_goals(mft, omft, N); _set_joint_goals(jt, ojt, q, 7)
    ctrl.updateControllerTaskModels()
    tau = ctrl.computeControlTorques()
    ref = ob.cycle()
    st = robot.status()
    assert (st & sp.capi.STATUS_UNHANDLED).sum() == 0
    assert ((st & sp.capi.STATUS_SINGULAR_PATH) == 0).mean() > 0.9
    assert rel_err(tau, ref).max() < REL_TOL


def test_config2_gravity_and_saturation():
    import sai_primitives_b200 as sp
    N = 64
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.075)
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    link, pt = TASK_POINTS["panda"]
    mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)))
    jt = sp.JointTask(robot)
    ctrl = sp.RobotController(robot, [mft, jt])
    ctrl.enableGravityCompensation(True); ctrl.enableTorqueSaturation(True)
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    omft = ob.add_mft(link, (np.eye(3), np.array(pt))); ojt = ob.add_jt()
    ob.finalize()
    for c in ob.controllers:
        c.enableGravityCompensation(True); c.enableTorqueSaturation(True)
    _set_mft_goals(mft, omft, N); _set_joint_goals(jt, ojt, q, 7)
    ctrl.updateControllerTaskModels()
    tau = ctrl.computeControlTorques()
    ref = ob.cycle()
    assert (robot.status() & sp.capi.STATUS_UNHANDLED).sum() == 0
    assert rel_err(tau, ref).max() < REL_TOL


def _states_near_limits(robot_name, N, seed_stream=11):
    """states with a few joints pushed into the position / velocity buffer zones (both zones, both directions)"""
    from oracle.robots import make_chain
    ch = make_chain(robot_name)
    q, dq, _ = sample_states(robot_name, N, min_sigma_ratio=0.075)
    expect_active = np.zeros(N, dtype=int)
    for i in range(N):
        g = rng_for(i, stream=seed_stream)
        for _ in range(int(g.integers(0, 4))):          # 0..3 joints per robot
            j = int(g.integers(0, ch.n)); kind = int(g.integers(0, 4)); depth = g.uniform(0.005, 0.15)
            if kind == 0: q[i, j] = ch.q_upper[j] - depth
            elif kind == 1: q[i, j] = ch.q_lower[j] + depth
            elif kind == 2: dq[i, j] = ch.dq_max[j] - 3.0 * depth
            else: dq[i, j] = -ch.dq_max[j] + 3.0 * depth
    return q, dq


@pytest.mark.parametrize("hier", ["jt", "mft_jt"])
@pytest.mark.parametrize("sat,grav", [(False, False), (True, True)])
def test_joint_limit_avoidance(hier, sat, grav):
    """RobotController::enableJointLimitAvoidance (RobotController.cpp:96-112) with the zone logic of
    JointLimitAvoidanceTask.cpp:174-421, on states inside the position and velocity buffer zones."""
    import sai_primitives_b200 as sp
    N = 96
    q, dq = _states_near_limits("panda", N)
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    link, pt = TASK_POINTS["panda"]
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    tasks = []
    if hier == "mft_jt":
        mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt))); tasks.append(mft)
        omft = ob.add_mft(link, (np.eye(3), np.array(pt)))
    jt = sp.JointTask(robot); tasks.append(jt)
    ojt = ob.add_jt()
    ctrl = sp.RobotController(robot, tasks)
    ctrl.enableJointLimitAvoidance(True); ctrl.enableGravityCompensation(grav); ctrl.enableTorqueSaturation(sat)
    ob.finalize()
    for c in ob.controllers:
        c.enableJointLimitAvoidance(True); c.enableGravityCompensation(grav); c.enableTorqueSaturation(sat)
    if hier == "mft_jt":
        _set_mft_goals(mft, omft, N)
    _set_joint_goals(jt, ojt, q, 7)
    ctrl.updateControllerTaskModels()
    tau = ctrl.computeControlTorques()
    ref = ob.cycle()
    active = np.array([c._joint_limit_avoidance_task._active_constraints for c in ob.controllers])
    assert (active > 0).sum() > N // 3 and (active == 0).sum() > 5 and active.max() >= 2
    st = robot.status()
    assert (st & sp.capi.STATUS_UNHANDLED).sum() == 0
    assert rel_err(tau, ref).max() < REL_TOL
    # switching it off again gives the plain controller back
    ctrl.enableJointLimitAvoidance(False)
    for c in ob.controllers:
        c.enableJointLimitAvoidance(False)
    ctrl.updateControllerTaskModels()
    assert rel_err(ctrl.computeControlTorques(), ob.cycle()).max() < REL_TOL


@pytest.mark.parametrize("dec", DEC)
@pytest.mark.parametrize("case", ["panda_xyz", "panda_yz_rotx", "rrrr_planar", "puma_full"])
def test_partial_and_other_robots(case, dec):
    """Partial MotionForceTasks (examples 08/09/11) and the other robots, JointTask in the null space."""
    import sai_primitives_b200 as sp
    cases = {
        "panda_xyz": ("panda", [(1, 0, 0), (0, 1, 0), (0, 0, 1)], []),              # examples/09:114-121
        "panda_yz_rotx": ("panda", [(0, 1, 0), (0, 0, 1)], [(1, 0, 0)]),             # examples/08:112-122
        "rrrr_planar": ("rrrr", [(1, 0, 0), (0, 1, 0)], [(0, 0, 1)]),                # examples/11:105-115
        "puma_full": ("puma_like", None, None),
    }
    robot_name, dt_, dr_ = cases[case]
    N = 64
    link, pt = TASK_POINTS[robot_name]
    B = None
    if dt_ is not None:
        cols = [np.concatenate([np.array(v, float), np.zeros(3)]) for v in dt_] + \
               [np.concatenate([np.zeros(3), np.array(v, float)]) for v in dr_]
        B = np.array(cols).T
    q, dq, _ = sample_states(robot_name, N, min_sigma_ratio=0.075, dirs=B)
    n = q.shape[1]
    robot = sp.BatchedRobot(robot_name, N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)), dt_, dr_)
    jt = sp.JointTask(robot)
    mft.setDynamicDecouplingType(dec); jt.setDynamicDecouplingType(dec)
    ctrl = sp.RobotController(robot, [mft, jt])
    ob = OracleBatch(robot_name, N); ob.set_state(q, dq)
    omft = ob.add_mft(link, (np.eye(3), np.array(pt)), dt_, dr_); ojt = ob.add_jt()
    for a, b in zip(omft, ojt):
        a.setDynamicDecouplingType(dec); b.setDynamicDecouplingType(dec)
    ob.finalize()
    _set_mft_goals(mft, omft, N); _set_joint_goals(jt, ojt, q, n)
    ctrl.updateControllerTaskModels()
    tau = ctrl.computeControlTorques()
    ref = ob.cycle()
    st = robot.status()
    assert (st & sp.capi.STATUS_UNHANDLED).sum() == 0
    assert ((st & sp.capi.STATUS_SINGULAR_PATH) == 0).mean() > 0.85
    assert rel_err(tau, ref).max() < REL_TOL
    if robot_name == "puma_like":
        assert ((st & sp.capi.STATUS_ZERO_RANGE) != 0).all()


@pytest.mark.parametrize("dec", DEC)
@pytest.mark.parametrize("handling", [True, False])
def test_config4_singular_path_multi_cycle(dec, handling):
    """BASELINE config 4: unfiltered Panda states (about half inside the reference's blending band
    6e-3 < s_i/s_0 < 6e-2) through the SVD path: classification, type-1/type-2 joint strategies and their
    memory (history, counters, q_prior) over several cycles with a moving state."""
    import sai_primitives_b200 as sp
    N = 96
    q, dq, _ = sample_states("panda", N)
    link, pt = TASK_POINTS["panda"]
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)))
    jt = sp.JointTask(robot)
    mft.setDynamicDecouplingType(dec); jt.setDynamicDecouplingType(dec)
    if not handling:
        mft.disableSingularityHandling()
    ctrl = sp.RobotController(robot, [mft, jt])
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    omft = ob.add_mft(link, (np.eye(3), np.array(pt))); ojt = ob.add_jt()
    for a, b in zip(omft, ojt):
        a.setDynamicDecouplingType(dec); b.setDynamicDecouplingType(dec)
        if not handling:
            a.disableSingularityHandling()
    ob.finalize()
    _set_mft_goals(mft, omft, N); _set_joint_goals(jt, ojt, q, 7)
    seen_t1 = seen_t2 = 0
    for cycle in range(4):
        ctrl.updateControllerTaskModels()
        tau = ctrl.computeControlTorques()
        ref = ob.cycle()
        st = robot.status()
        assert (st & sp.capi.STATUS_UNHANDLED).sum() == 0
        sing = np.array([len(t._singularity_handler._singularity_types) != 0 for t in omft])
        assert ((st & sp.capi.STATUS_SINGULAR_PATH) != 0)[sing].all()
        assert rel_err(tau, ref).max() < REL_TOL, cycle
        seen_t1 += int(((st & sp.capi.STATUS_TYPE1) != 0).sum()); seen_t2 += int(((st & sp.capi.STATUS_TYPE2) != 0).sum())
        # move every robot a little (semi-implicit Euler with the commanded torque direction is not needed: any motion will do)
        q = q + 0.002 * dq
        robot.setQ(q); robot.updateModel(); ob.set_state(q, dq)
    assert sing.sum() > 10
    if handling and dec != 2:
        assert seen_t1 > 0 and seen_t2 > 0     # both joint strategies were exercised


def test_singular_path_other_robots_and_partial_tasks():
    """PUMA-like 6R (wrist/elbow alignments) and the planar 4R arm (stretched) near their singularities."""
    import sai_primitives_b200 as sp
    for name, dt_, dr_ in (("puma_like", None, None), ("rrrr", [(1, 0, 0), (0, 1, 0)], [(0, 0, 1)]),
                           ("panda", [(0, 1, 0), (0, 0, 1)], [(1, 0, 0)])):
        N = 64
        q, dq, _ = sample_states(name, N)
        n = q.shape[1]
        for i in range(0, N, 4):          # push a quarter of the robots towards a kinematic singularity
            if name == "puma_like":
                q[i, 4] = 0.01 * (1 + i % 3)
            elif name == "rrrr":
                q[i, 1:] = 0.02 * (1 + i % 3)
            else:
                q[i, 3] = -0.08
        link, pt = TASK_POINTS[name]
        robot = sp.BatchedRobot(name, N)
        robot.setQ(q); robot.setDq(dq); robot.updateModel()
        mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)), dt_, dr_)
        jt = sp.JointTask(robot)
        ctrl = sp.RobotController(robot, [mft, jt])
        ob = OracleBatch(name, N); ob.set_state(q, dq)
        omft = ob.add_mft(link, (np.eye(3), np.array(pt)), dt_, dr_); ojt = ob.add_jt(); ob.finalize()
        _set_mft_goals(mft, omft, N); _set_joint_goals(jt, ojt, q, n)
        for cycle in range(3):
            ctrl.updateControllerTaskModels()
            tau = ctrl.computeControlTorques()
            ref = ob.cycle()
            assert (robot.status() & sp.capi.STATUS_UNHANDLED).sum() == 0
            assert rel_err(tau, ref).max() < REL_TOL, (name, cycle)
        sing = np.array([len(t._singularity_handler._singularity_types) != 0 for t in omft])
        assert sing.sum() >= (N // 8 if name != "panda" else 1), name


def test_general_hierarchies():
    """Hierarchies without a specialised kernel run on the general-hierarchy kernel:
    ex.06 (partial JointTask on slider + last joint, then a 6-DoF MotionForceTask) on the 8-DoF sliding-base Panda,
    and [partial MotionForceTask (position), partial MotionForceTask (orientation), JointTask] on the Panda."""
    import sai_primitives_b200 as sp
    # ---- examples/06-partial_joint_task/06-partial_joint_task.cpp:108-128
    name = "panda_sliding_base"
    N = 48
    q, dq, _ = sample_states(name, N)
    link, pt = TASK_POINTS[name]
    S = np.zeros((2, 8)); S[0, 0] = 1; S[1, 7] = 1
    robot = sp.BatchedRobot(name, N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    pjt = sp.JointTask(robot, S, task_name="partial_joint_task")
    mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)))
    ctrl = sp.RobotController(robot, [pjt, mft])
    ob = OracleBatch(name, N); ob.set_state(q, dq)
    opjt = ob.add_jt(S, name="partial_joint_task"); omft = ob.add_mft(link, (np.eye(3), np.array(pt))); ob.finalize()
    _set_mft_goals(mft, omft, N)
    gp = np.array([S @ q[i] + rng_for(i, stream=5).uniform(-0.2, 0.2, 2) for i in range(N)])
    pjt.setGoalPosition(gp)
    for i in range(N):
        opjt[i].setGoalPosition(gp[i])
    for cycle in range(3):
        ctrl.updateControllerTaskModels()
        tau = ctrl.computeControlTorques()
        ref = ob.cycle()
        assert (robot.status() & sp.capi.STATUS_UNHANDLED).sum() == 0
        assert rel_err(tau, ref).max() < REL_TOL, cycle
    # ---- two partial motion-force tasks + joint task
    N = 48
    q, dq, _ = sample_states("panda", N)
    link, pt = TASK_POINTS["panda"]
    xyz = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    t1 = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)), xyz, [], task_name="position")
    t2 = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)), [], xyz, task_name="orientation")
    jt = sp.JointTask(robot)
    ctrl = sp.RobotController(robot, [t1, t2, jt])
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    o1 = ob.add_mft(link, (np.eye(3), np.array(pt)), xyz, [], name="position")
    o2 = ob.add_mft(link, (np.eye(3), np.array(pt)), [], xyz, name="orientation")
    oj = ob.add_jt(); ob.finalize()
    _set_mft_goals(t1, o1, N); _set_mft_goals(t2, o2, N); _set_joint_goals(jt, oj, q, 7)
    for cycle in range(3):
        ctrl.updateControllerTaskModels()
        tau = ctrl.computeControlTorques()
        ref = ob.cycle()
        assert rel_err(tau, ref).max() < REL_TOL, cycle
    # a task after a full joint task is rejected like RobotController.cpp:45-51
    robot = sp.BatchedRobot("panda", 4)
    a = sp.JointTask(robot, task_name="a"); b = sp.JointTask(robot, task_name="b")
    with pytest.raises(ValueError):
        sp.RobotController(robot, [a, b])


def test_config4_mixed_dof_batch():
    """BASELINE config 4: one batch mixing Panda (7), PUMA-like (6) and planar RRRR (4) robots, grouped by model on
    the host (sharding.ModelGroups), unfiltered states (singular robots included), two cycles."""
    import sai_primitives_b200 as sp
    from sai_primitives_b200.sharding import ModelGroups
    specs = {"panda": (None, None), "puma_like": (None, None), "rrrr": ([(1, 0, 0), (0, 1, 0)], [(0, 0, 1)])}
    N = 90
    models = [("panda", "puma_like", "rrrr")[i % 3] for i in range(N)]

    class GpuGroup:
        def __init__(self, name, n):
            self.name, self.n = name, n
            self.robot = sp.BatchedRobot(name, n)
            self.ctrl = None
        def set_state(self, q, dq):
            self.robot.setQ(q); self.robot.setDq(dq); self.robot.updateModel()
            if self.ctrl is None:
                link, pt = TASK_POINTS[self.name]
                dt_, dr_ = specs[self.name]
                self.mft = sp.MotionForceTask(self.robot, link, (np.eye(3), np.array(pt)), dt_, dr_)
                self.jt = sp.JointTask(self.robot)
                self.ctrl = sp.RobotController(self.robot, [self.mft, self.jt])
        def cycle(self):
            self.ctrl.updateControllerTaskModels()
            return self.ctrl.computeControlTorques()

    class OracleGroup:
        def __init__(self, name, n):
            self.name, self.n = name, n
            self.ob = OracleBatch(name, n)
            self.ready = False
        def set_state(self, q, dq):
            self.ob.set_state(q, dq)
            if not self.ready:
                link, pt = TASK_POINTS[self.name]
                dt_, dr_ = specs[self.name]
                self.ob.add_mft(link, (np.eye(3), np.array(pt)), dt_, dr_); self.ob.add_jt(); self.ob.finalize()
                self.ready = True
        def cycle(self):
            return self.ob.cycle()

    per_model = {nm: sample_states(nm, N)[:2] for nm in specs}
    q = [per_model[m][0][i] for i, m in enumerate(models)]
    dq = [per_model[m][1][i] for i, m in enumerate(models)]
    gpu = ModelGroups(models, GpuGroup); ora = ModelGroups(models, OracleGroup)
    gpu.set_state(q, dq); ora.set_state(q, dq)
    for cycle in range(2):
        a, b = gpu.cycle(), ora.cycle()
        for i in range(N):
            assert a[i].shape == b[i].shape
            assert np.abs(a[i] - b[i]).max() <= REL_TOL * max(np.abs(b[i]).max(), 1e-9), (cycle, i, models[i])
        q = [qi + 0.002 * dqi for qi, dqi in zip(q, dq)]
        gpu.set_state(q, dq); ora.set_state(q, dq)


def test_config3_surface_contact_moment_control_in_compliant_frame():
    """BASELINE config 3, second variant (examples/07-surface_surface_contact/...cpp:135-201): full 6-DoF task with the
    force/motion spaces parametrised in the compliant frame, force space dim 1 + moment space dim 2 about local Z,
    closed-loop force (with passivity) and closed-loop moment, custom force/moment gains, sensor frame set,
    MotionForceTask alone in the controller; 120 cycles with a moving state and noisy sensed wrench."""
    import sai_primitives_b200 as sp
    N = 24
    K = 120
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.08)
    link, pt = TASK_POINTS["panda"]
    comp = (np.eye(3), np.array(pt))
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    mft = sp.MotionForceTask(robot, link, comp, task_name="surface_alignment_task",
                             is_force_motion_parametrization_in_compliant_frame=True)
    mft.enablePassivity()
    sensor_R = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]]); sensor_t = np.array([0.01, 0.02, 0.05])
    mft.setForceSensorFrame(link, (sensor_R, sensor_t))
    ctrl = sp.RobotController(robot, [mft])
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    omft = ob.add_mft(link, comp, in_compliant=True, name="surface_alignment_task"); ob.finalize()
    for t in omft:
        t.enablePassivity(); t.setForceSensorFrame(link, (sensor_R, sensor_t))
    _set_mft_goals(mft, omft, N)

    def sensed(k):
        f = np.zeros((N, 3)); m = np.zeros((N, 3))
        for i in range(N):
            g = rng_for(i * 7919 + k, stream=8)
            f[i] = np.array([0.3, -0.2, 9.0]) + g.normal(0, 1.5, 3); m[i] = g.normal(0, 0.3, 3)
        return f, m

    for k in range(K):
        if k == 5:   # contact detected: switch spaces, close the loops (examples/07:188-201)
            assert mft.parametrizeForceMotionSpaces(1, (0, 0, 1)) is True
            assert mft.parametrizeMomentRotMotionSpaces(2, (0, 0, 1)) is True
            mft.setClosedLoopForceControl(); mft.setClosedLoopMomentControl()
            mft.setGoalForce(np.array([0, 0, 10.0])); mft.setGoalMoment(np.zeros(3))
            mft.setForceControlGains(0.7, 5.0, 1.5); mft.setMomentControlGains(0.7, 4.0, 1.5)
            for t in omft:
                t.parametrizeForceMotionSpaces(1, (0, 0, 1)); t.parametrizeMomentRotMotionSpaces(2, (0, 0, 1))
                t.setClosedLoopForceControl(); t.setClosedLoopMomentControl()
                t.setGoalForce((0, 0, 10.0)); t.setGoalMoment((0, 0, 0))
                t.setForceControlGains(0.7, 5.0, 1.5); t.setMomentControlGains(0.7, 4.0, 1.5)
        f, m = sensed(k)
        mft.updateSensedForceAndMoment(f, m)
        for i in range(N):
            omft[i].updateSensedForceAndMoment(f[i], m[i])
        ctrl.updateControllerTaskModels()
        tau = ctrl.computeControlTorques()
        ref = ob.cycle()
        assert (robot.status() & sp.capi.STATUS_UNHANDLED).sum() == 0
        assert rel_err(tau, ref).max() < REL_TOL, k
        q = q + 0.001 * dq
        robot.setQ(q); robot.updateModel(); ob.set_state(q, dq)
    sf = mft.getSensedForceControlWorldFrame()
    assert np.abs(sf - np.array([t._sensed_force_control_world_frame for t in omft])).max() < 1e-12


def test_velocity_saturation_and_anisotropic_gains():
    """Velocity saturation in both tasks (MotionForceTask.cpp:416-461, JointTask.cpp:327-340) and per-axis gains."""
    import sai_primitives_b200 as sp
    N = 48
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.08)
    link, pt = TASK_POINTS["panda"]
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt)))
    jt = sp.JointTask(robot)
    mft.enableVelocitySaturation(0.05, 0.2); mft.setPosControlGains([100, 150, 80], [20, 25, 15], [2, 0, 1]); mft.setOriControlGains(150, 25, 3)
    jt.enableVelocitySaturation(0.3); jt.setGains(np.linspace(40, 70, 7), np.linspace(10, 16, 7), np.linspace(0, 3, 7))
    ctrl = sp.RobotController(robot, [mft, jt])
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    omft = ob.add_mft(link, (np.eye(3), np.array(pt))); ojt = ob.add_jt(); ob.finalize()
    for a, b in zip(omft, ojt):
        a.enableVelocitySaturation(0.05, 0.2); a.setPosControlGains([100, 150, 80], [20, 25, 15], [2, 0, 1]); a.setOriControlGains(150, 25, 3)
        b.enableVelocitySaturation(0.3); b.setGains(np.linspace(40, 70, 7), np.linspace(10, 16, 7), np.linspace(0, 3, 7))
    _set_mft_goals(mft, omft, N); _set_joint_goals(jt, ojt, q, 7, with_vel=True)
    for cycle in range(3):
        ctrl.updateControllerTaskModels()
        assert rel_err(ctrl.computeControlTorques(), ob.cycle()).max() < REL_TOL
    with pytest.raises(ValueError):
        mft.setPosControlGains(-1.0, 20.0)          # MotionForceTask.cpp:583-587
    with pytest.raises(ValueError):
        jt.enableVelocitySaturation(-0.1)             # JointTask.cpp:409-413
    with pytest.raises(NotImplementedError):
        jt.enableInternalOtgJerkLimited(1.0, 2.0, 3.0)      # only the acceleration-limited generator is built (row f-4)


def test_full_size_properties_262144_robots():
    """BASELINE config 3 size (262,144 robots): size-independent properties instead of an oracle loop --
    (i) the batch result equals the same robots evaluated in shuffled order and in chunks (no cross-robot coupling,
    no dependence on the batch index), (ii) duplicated robots get bit-identical torques, (iii) an 8,192-robot
    sample matches the C++ oracle."""
    import sai_primitives_b200 as sp
    from oracle.cpp_ref import CppOracleBatch
    N = 262144
    base_q, base_dq, _ = sample_states("panda", 2048, min_sigma_ratio=0.075)
    rng = np.random.default_rng(77)
    pick = rng.integers(0, 2048, N)
    q, dq = base_q[pick], base_dq[pick]
    link, pt = TASK_POINTS["panda"]

    def run(qs, dqs):
        robot = sp.BatchedRobot("panda", qs.shape[0])
        robot.setQ(qs); robot.setDq(dqs); robot.updateModel()
        mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt))); jt = sp.JointTask(robot)
        ctrl = sp.RobotController(robot, [mft, jt])
        mft.setGoalLinearVelocity(np.array([0.02, -0.01, 0.03])); jt.setGoalPosition(qs + 0.05)
        ctrl.updateControllerTaskModels()
        tau = ctrl.computeControlTorques()
        st = robot.status()
        robot.close()
        return tau, st
    tau, st = run(q, dq)
    assert (st & sp.capi.STATUS_UNHANDLED).sum() == 0 and np.isfinite(tau).all()
    # (ii) duplicates are bit-identical
    first = {}
    for i in range(0, N, 997):
        first.setdefault(int(pick[i]), i)
        assert np.array_equal(tau[i], tau[first[int(pick[i])]])
    # (i) shuffled order
    perm = rng.permutation(N)
    tau_p, _ = run(q[perm], dq[perm])
    assert np.array_equal(tau_p, tau[perm])
    # (iii) sample against the C++ oracle
    idx = np.arange(0, N, 32)
    cb = CppOracleBatch("panda", idx.size); cb.set_state(q[idx], dq[idx])
    tm = cb.add_mft(link, (np.eye(3), np.array(pt))); tj = cb.add_jt()
    x0, R0 = cb.mft_get_current(tm)
    z = np.zeros((idx.size, 3))
    cb.mft_set_goals(tm, x0, R0, np.tile([0.02, -0.01, 0.03], (idx.size, 1)), z, z, z); cb.jt_set_goals(tj, q[idx] + 0.05)
    ref = cb.cycle(n_threads=8)
    assert rel_err(tau[idx], ref).max() < REL_TOL


@pytest.mark.parametrize("filtered", [False, True])
def test_pipelined_back_to_back_cycles(filtered):
    """Cross-cycle pipelining (csrc/osc_pipeline.cuh): K cycles enqueued back to back, so that blocks of cycle c + 1 run while
    the last wave of cycle c is still in flight, with integral gains on both tasks (every cycle reads what the previous one
    wrote) and unfiltered states (hand-overs to the general path in most blocks, i.e. the dirty-block wait).  The result must be
    bit-identical to the same cycles run one by one without pipelining, and match the oracle on a sample."""
    import os
    import torch
    import sai_primitives_b200 as sp
    # filtered: nothing is handed over, every block of cycle c + 1 only waits for its own predecessor (the steady state of the
    # benchmark, many cycles deep); unfiltered: hand-overs in most blocks
    N, K = (66000, 60) if filtered else (66000, 10)
    base_q, base_dq, _ = sample_states("panda", 1024, min_sigma_ratio=0.1 if filtered else None)
    rng = np.random.default_rng(5)
    pick = rng.integers(0, 1024, N)
    pick[:64] = np.arange(64)
    q, dq = base_q[pick], base_dq[pick]
    link, pt = TASK_POINTS["panda"]
    dev = torch.device("cuda", 0)

    def build(pipelined):
        if not pipelined:
            os.environ["SAI_B200_NO_PIPELINE"] = "1"
        try:
            robot = sp.BatchedRobot("panda", N)
        finally:
            os.environ.pop("SAI_B200_NO_PIPELINE", None)
        robot.setQ(q); robot.setDq(dq); robot.updateModel()
        mft = sp.MotionForceTask(robot, link, (np.eye(3), np.array(pt))); jt = sp.JointTask(robot)
        mft.setPosControlGains(100.0, 20.0, 5.0); mft.setOriControlGains(200.0, 28.3, 4.0); jt.setGains(50.0, 14.0, 3.0)
        ctrl = sp.RobotController(robot, [mft, jt])
        mft.setGoalLinearVelocity(np.array([0.02, -0.01, 0.03])); jt.setGoalPosition(q + 0.05)
        return robot, mft, jt, ctrl

    d_q = torch.from_numpy(np.ascontiguousarray(q.T)).to(dev); d_dq = torch.from_numpy(np.ascontiguousarray(dq.T)).to(dev)
    out = {}
    for pipelined in (True, False):
        robot, mft, jt, ctrl = build(pipelined)
        d_tau = torch.zeros((7, N), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        for k in range(K):
            ctrl.stepDevice(d_q.data_ptr(), d_dq.data_ptr(), d_tau.data_ptr())
            if not pipelined:
                robot.sync()
        robot.sync()
        st = robot.status()
        assert (st & sp.capi.STATUS_UNHANDLED).sum() == 0
        out[pipelined] = (d_tau.cpu().numpy().T.copy(), mft._get(sp.capi.MFT_INTEGRATED_POSITION_ERROR), jt._get(sp.capi.JT_INTEGRATED_POSITION_ERROR), st)
        robot.close()
    if filtered:
        assert ((out[True][3] & sp.capi.STATUS_SINGULAR_PATH) != 0).sum() == 0
    else:
        assert ((out[True][3] & sp.capi.STATUS_SINGULAR_PATH) != 0).mean() > 0.3      # most blocks handed robots over
    if os.environ.get("SAI_B200_BLEND_SPLIT") == "1" and not filtered:
        # the split blending path (forced on for the pipelined handle) starts from the fused kernel's kinematics, the single
        # blending kernel of the other handle from its own: same results to rounding, not bit for bit
        for a, b in zip(out[True][:3], out[False][:3]):
            assert np.abs(a - b).max() <= 1e-9 * max(1.0, np.abs(b).max())
    else:
        for a, b in zip(out[True][:3], out[False][:3]):
            assert np.array_equal(a, b)
    ob = OracleBatch("panda", 64); ob.set_state(q[:64], dq[:64])
    omft = ob.add_mft(link, (np.eye(3), np.array(pt))); ojt = ob.add_jt(); ob.finalize()
    for i in range(64):
        omft[i].setPosControlGains(100.0, 20.0, 5.0); omft[i].setOriControlGains(200.0, 28.3, 4.0); ojt[i].setGains(50.0, 14.0, 3.0)
        omft[i].setGoalLinearVelocity((0.02, -0.01, 0.03)); ojt[i].setGoalPosition(q[i] + 0.05)
    for k in range(K):
        ref = ob.cycle()
    assert rel_err(out[True][0][:64], ref).max() < REL_TOL


@pytest.mark.parametrize("case", ["panda_full", "panda_xyz_force", "sliding_partial_joint"])
def test_task_observers_and_nullspaces(case):
    """TemplateTask::getTaskNullspace / getPreviousTasksNullspace / getTaskAndPreviousNullspace (TemplateTask.h:74-89),
    MotionForceTask::getPositionError / getOrientationError / sigma*() (MotionForceTask.h:268-269, :613-616) and the per-cycle
    observers getCurrentLinearVelocity / getCurrentAngularVelocity / getUnitMassForce, against the reference's own getters."""
    import sai_primitives_b200 as sp
    N = 40
    name = "panda_sliding_base" if case == "sliding_partial_joint" else "panda"
    q, dq, _ = sample_states(name, N)          # unfiltered: singular robots included (their N goes through the blending branch)
    n = q.shape[1]
    link, pt = TASK_POINTS[name]
    comp = (np.eye(3), np.array(pt))
    robot = sp.BatchedRobot(name, N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    ob = OracleBatch(name, N); ob.set_state(q, dq)
    if case == "panda_full":
        gt = [sp.MotionForceTask(robot, link, comp), sp.JointTask(robot)]
        ot = [ob.add_mft(link, comp), ob.add_jt()]
    elif case == "panda_xyz_force":
        xyz = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
        gt = [sp.MotionForceTask(robot, link, comp, xyz, [], is_force_motion_parametrization_in_compliant_frame=True), sp.JointTask(robot)]
        ot = [ob.add_mft(link, comp, xyz, [], in_compliant=True), ob.add_jt()]
        gt[0].parametrizeForceMotionSpaces(1, (0, 0, 1))
        for t in ot[0]:
            t.parametrizeForceMotionSpaces(1, (0, 0, 1))
    else:
        S = np.zeros((2, 8)); S[0, 0] = 1; S[1, 7] = 1
        gt = [sp.JointTask(robot, S, task_name="partial_joint_task"), sp.MotionForceTask(robot, link, comp)]
        ot = [ob.add_jt(S, name="partial_joint_task"), ob.add_mft(link, comp)]
    ctrl = sp.RobotController(robot, gt); ob.finalize()
    mi = 1 if case == "sliding_partial_joint" else 0
    _set_mft_goals(gt[mi], ot[mi], N)
    ctrl.updateControllerTaskModels()
    tau = ctrl.computeControlTorques()
    assert rel_err(tau, ob.cycle()).max() < REL_TOL
    for g, o in zip(gt, ot):
        for name_ in ("getTaskNullspace", "getPreviousTasksNullspace", "getTaskAndPreviousNullspace"):
            a = getattr(g, name_)()
            b = np.array([np.asarray(getattr(t, name_)()).reshape(n, n) for t in o])
            assert a.shape == (N, n, n)
            assert np.abs(a - b).max() < 1e-8 * max(1.0, np.abs(b).max()), (case, name_)
    g, o = gt[mi], ot[mi]
    for name_ in ("getPositionError", "getOrientationError", "getCurrentLinearVelocity", "getCurrentAngularVelocity", "getUnitMassForce",
                  "sigmaForce", "sigmaPosition", "sigmaMoment", "sigmaOrientation"):
        a = getattr(g, name_)()
        b = np.array([np.asarray(getattr(t, name_)()) for t in o]).reshape(a.shape)
        assert np.abs(a - b).max() < 1e-9 * max(1.0, np.abs(b).max()), (case, name_)
    # observers can be switched off; the fields then refuse to answer instead of returning stale values
    robot.enableObservers(False)
    with pytest.raises(Exception):
        g.getUnitMassForce()
    robot.enableObservers(True)


def test_config3_full_size_262144_robots_sampled_parity():
    """BASELINE config 3 at its full size: 262,144 Pandas, partial MotionForceTask (XYZ) with force space
    dimension 1 about Z, closed-loop force + POPC passivity, JointTask in the null space; 60 consecutive cycles with a noisy
    sensed force and a drifting state.  Properties that do not need an oracle loop over the whole batch -- duplicated robots get
    bit-identical torques and POPC state, no robot leaves the CUDA path -- plus a 2,048-robot sample against the CPU checker
    (the C++ port, itself checked against the reference's compiled control law) at cycles 1, 50, 51 and 60."""
    import sai_primitives_b200 as sp
    from oracle.cpp_ref import CppOracleBatch
    N, K, NS = 262144, 60, 2048
    dirs = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
    base_q, base_dq, _ = sample_states("panda", NS, min_sigma_ratio=0.1, dirs=np.eye(6)[:, :3])
    rng = np.random.default_rng(12)
    pick = rng.integers(0, NS, N); pick[:NS] = np.arange(NS)
    q0, dq = base_q[pick], base_dq[pick]
    link, pt = TASK_POINTS["panda"]
    comp = (np.eye(3), np.array(pt))
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(q0); robot.setDq(dq); robot.updateModel()
    mft = sp.MotionForceTask(robot, link, comp, dirs, []); jt = sp.JointTask(robot)
    ctrl = sp.RobotController(robot, [mft, jt])
    assert mft.parametrizeForceMotionSpaces(1, (0, 0, 1)) is True
    mft.setGoalForce(np.array([0, 0, -5.0])); mft.setClosedLoopForceControl(); mft.enablePassivity()
    cb = CppOracleBatch("panda", NS); cb.set_state(base_q, base_dq)
    tm = cb.add_mft(link, comp, dirs, []); cb.add_jt()
    cb.mft_force_setup(tm, fdim=1, faxis=(0, 0, 1), cl_force=True, passivity=True)
    cb.mft_set_force_goals(tm, np.tile([0, 0, -5.0], (NS, 1)), np.zeros((NS, 3)))
    frng = np.random.default_rng(13)
    for k in range(1, K + 1):
        fs = np.array([0, 0, -5.0]) + frng.normal(0, 2.0, (NS, 3)); ms = frng.normal(0, 0.1, (NS, 3))
        mft.updateSensedForceAndMoment(fs[pick], ms[pick]); cb.mft_update_sensed(tm, fs, ms)
        ctrl.updateControllerTaskModels()
        tau = ctrl.computeControlTorques()
        ref = cb.cycle(n_threads=8)
        if k in (1, 50, 51, K):
            assert rel_err(tau[:NS], ref).max() < REL_TOL, k
        qk = base_q + 0.0005 * k * base_dq
        robot.setQ(qk[pick]); robot.updateModel(); cb.set_state(qk, base_dq)
    st = robot.status()
    assert (st & (sp.capi.STATUS_UNHANDLED | sp.capi.STATUS_POPC_OVERFLOW)).sum() == 0 and np.isfinite(tau).all()
    popc = mft._get(sp.capi.MFT_POPC_STATE)
    first = {}
    for i in range(NS, N, 499):            # duplicates of a sampled robot: bit-identical torques and passivity state
        j = int(pick[i])
        assert np.array_equal(tau[i], tau[j]) and np.array_equal(popc[i], popc[j])


def test_split_blending_path_131072_robots_sampled_parity():
    """Config 4's branch at scale: 131,072 uniformly sampled Panda states, 57 % of them inside the reference's blending band
    (SingularityHandler.cpp:75-368).  The first cycle runs them through the single blending kernel; its hand-over count tells the
    host that the list is longer than that kernel holds at once, and the following cycles take the split path (the fused kernel
    parks kinematics and dynamics, classification kernel, one variant per warp, general path for the rest).  A 256-robot sample
    against the reference's compiled control law at every cycle, duplicates bit-identical, and the distribution over the variants
    from the debug counters: the blending kernels, not the rolled general path, must have handled the robots."""
    import ctypes as C
    import os
    import sai_primitives_b200 as sp
    N, NS, K = 131072, 256, 4
    base_q, base_dq, _ = sample_states("panda", NS)
    rng = np.random.default_rng(21)
    pick = rng.integers(0, NS, N); pick[:NS] = np.arange(NS)
    link, pt = TASK_POINTS["panda"]
    comp = (np.eye(3), np.array(pt))
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(base_q[pick]); robot.setDq(base_dq[pick]); robot.updateModel()
    mft = sp.MotionForceTask(robot, link, comp); jt = sp.JointTask(robot)
    ctrl = sp.RobotController(robot, [mft, jt])
    ob = OracleBatch("panda", NS); ob.set_state(base_q, base_dq)
    omft = ob.add_mft(link, comp); ojt = ob.add_jt()
    ob.finalize()
    x0 = mft.getCurrentPosition()[:NS]; R0 = mft.getCurrentOrientation()[:NS]
    xd = np.zeros((NS, 3)); Rd = np.zeros((NS, 3, 3)); gq = np.zeros((NS, 7))
    for i in range(NS):
        g = rng_for(i, stream=5)
        xd[i] = x0[i] + g.uniform(-0.05, 0.05, 3); Rd[i] = R0[i] @ rot_exp(g.uniform(-0.2, 0.2, 3)); gq[i] = base_q[i] + g.uniform(-0.2, 0.2, 7)
        omft[i].setGoalPosition(xd[i]); omft[i].setGoalOrientation(Rd[i]); ojt[i].setGoalPosition(gq[i])
    mft.setGoalPosition(xd[pick]); mft.setGoalOrientation(Rd[pick]); jt.setGoalPosition(gq[pick])
    lib = sp.capi.load_library()
    counts = (C.c_int32 * 4)()
    for k in range(K):
        ctrl.updateControllerTaskModels()
        tau = ctrl.computeControlTorques()
        ref = ob.cycle()
        assert rel_err(tau[:NS], ref).max() < REL_TOL, k
        st = robot.status()
        assert (st & sp.capi.STATUS_UNHANDLED).sum() == 0
        assert lib.osc_debug_general_path_counts(robot.handle, counts) == 0
        if k == 0 and os.environ.get("SAI_B200_BLEND_SPLIT") != "1":
            assert list(counts) == [-1, -1, -1, -1]          # the single kernel took the first cycle
        else:
            c = np.array(list(counts))
            on_general_path = int(((st & sp.capi.STATUS_SINGULAR_PATH) != 0).sum())
            assert c[1] + c[2] > 0.4 * N and c[3] < 0.02 * N, c
            assert on_general_path <= c.sum() <= on_general_path + c[0] + c[3]    # variant 0 / general path: may turn out non-singular
        qk = base_q + 0.002 * (k + 1) * base_dq
        robot.setQ(qk[pick]); robot.updateModel(); ob.set_state(qk, base_dq)
    for i in range(NS, N, 997):
        assert np.array_equal(tau[i], tau[int(pick[i])])
