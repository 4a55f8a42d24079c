"""The numpy restatement (oracle/primitives.py) and the C++ port (oracle/cpp/osc_ref.cpp) against THE REFERENCE'S OWN
CONTROL LAW: oracle/_ref/libsai_ref_orient.so is /root/reference/src/{RobotController,tasks/JointTask,tasks/MotionForceTask,
tasks/SingularityHandler,tasks/JointLimitAvoidanceTask,helper_modules/POPCExplicitForceControl,...}.cpp compiled where they
lie, unmodified (oracle/Makefile; stand-ins for Eigen and sai-model only).  Every scenario family of the GPU parity tests is
replayed here on the CPU: SURVEY.md rows a1-a16, f-2.  Runs wherever the library is present (it is built in this container)."""
import numpy as np
import pytest

from oracle import sai_ref
from tests.osc_testlib import TASK_POINTS, OracleBatch, rel_err, rng_for, rot_exp, sample_states

pytestmark = pytest.mark.skipif(not sai_ref.available(oriented=True), reason="oracle/_ref/libsai_ref_orient.so needs /root/reference to be built")

TOL = 1e-10   # two CPU implementations of the same FP64 arithmetic: far inside the 1e-9 budget of the GPU tests
XYZ = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]


def both(robot_name, N, **kw):
    return OracleBatch(robot_name, N, kind="numpy", **kw), OracleBatch(robot_name, N, kind="reference", **kw)


def set_goals(omft, ojt, q, stream=2, with_vel=False):
    N, n = q.shape
    for i in range(N):
        g = rng_for(i, stream=stream)
        if omft is not None:
            t = omft[i]
            x0, R0 = np.array(t._current_position), np.array(t._current_orientation)
            t.setGoalPosition(x0 + g.uniform(-0.05, 0.05, 3)); t.setGoalOrientation(R0 @ rot_exp(g.uniform(-0.2, 0.2, 3)))
            t.setGoalLinearVelocity(g.uniform(-0.1, 0.1, 3)); t.setGoalAngularVelocity(g.uniform(-0.1, 0.1, 3))
            t.setGoalLinearAcceleration(g.uniform(-0.5, 0.5, 3)); t.setGoalAngularAcceleration(g.uniform(-0.5, 0.5, 3))
        if ojt is not None:
            k = len(ojt[i].getCurrentPosition())
            g2 = rng_for(i, stream=stream + 100)
            ojt[i].setGoalPosition(np.array(ojt[i].getCurrentPosition()) + g2.uniform(-0.2, 0.2, k))
            if with_vel:
                ojt[i].setGoalVelocity(g2.uniform(-0.1, 0.1, k)); ojt[i].setGoalAcceleration(g2.uniform(-0.5, 0.5, k))


def compare_cycles(a, b, q, dq, cycles, use_prev=True, move=0.002):
    worst = 0.0
    for c in range(cycles):
        ta, tb = a.cycle(use_prev=use_prev), b.cycle(use_prev=use_prev)
        worst = max(worst, rel_err(ta, tb).max())
        assert worst < TOL, (c, worst)
        if move:
            q = q + move * dq
            a.set_state(q, dq); b.set_state(q, dq)
    return worst


@pytest.mark.parametrize("dec", [0, 1, 2])
@pytest.mark.parametrize("robot_name", ["panda", "rrrr", "panda_sliding_base"])
def test_config1_joint_task(robot_name, dec):
    N = 12
    q, dq, _ = sample_states(robot_name, N)
    a, b = both(robot_name, N)
    for ob in (a, b):
        ob.set_state(q, dq)
        jt = ob.add_jt(); ob.finalize()
        for t in jt:
            t.setGains(100.0, 20.0, 3.0); t.setDynamicDecouplingType(dec)
        set_goals(None, jt, q, with_vel=True)
    compare_cycles(a, b, q, dq, 3)


@pytest.mark.parametrize("use_prev", [True, False])
@pytest.mark.parametrize("handling", [True, False])
@pytest.mark.parametrize("dec", [0, 1, 2])
def test_config2_and_4_osc_with_singular_states(dec, handling, use_prev):
    """unfiltered Panda states: about half take the blending branch; 5 moving cycles so that the handler memory evolves"""
    N = 40
    q, dq, _ = sample_states("panda", N)
    link, pt = TASK_POINTS["panda"]
    a, b = both("panda", N)
    tasks = []
    for ob in (a, b):
        ob.set_state(q, dq)
        mft = ob.add_mft(link, (np.eye(3), np.array(pt))); jt = ob.add_jt(); ob.finalize()
        for x, y in zip(mft, jt):
            x.setDynamicDecouplingType(dec); y.setDynamicDecouplingType(dec)
            if not handling:
                x.disableSingularityHandling()
        set_goals(mft, jt, q)
        tasks.append(mft)
    compare_cycles(a, b, q, dq, 5, use_prev=use_prev)
    ta = [list(map(int, t._singularity_handler._singularity_types)) for t in tasks[0]]
    tb = [t._singularity_handler._singularity_types for t in tasks[1]]
    assert ta == tb
    assert sum(1 for t in ta if t) > 8
    for x, y in zip(tasks[0], tasks[1]):
        assert x._singularity_handler._type_1_counter == y._singularity_handler._type_1_counter
        assert x._singularity_handler._type_2_counter == y._singularity_handler._type_2_counter


@pytest.mark.parametrize("case", ["panda_xyz", "panda_yz_rotx", "rrrr_planar", "puma_full"])
def test_partial_tasks_and_other_robots(case):
    cases = {"panda_xyz": ("panda", XYZ, []), "panda_yz_rotx": ("panda", [(0, 1, 0), (0, 0, 1)], [(1, 0, 0)]),
             "rrrr_planar": ("rrrr", [(1, 0, 0), (0, 1, 0)], [(0, 0, 1)]), "puma_full": ("puma_like", None, None)}
    name, dt_, dr_ = cases[case]
    N = 24
    q, dq, _ = sample_states(name, N)
    for i in range(0, N, 4):          # push a quarter of the robots towards a kinematic singularity
        if name == "puma_like":
            q[i, 4] = 0.01 * (1 + i % 3)
        elif name == "rrrr":
            q[i, 1:] = 0.02 * (1 + i % 3)
        else:
            q[i, 3] = -0.08
    link, pt = TASK_POINTS[name]
    a, b = both(name, N)
    for ob in (a, b):
        ob.set_state(q, dq)
        mft = ob.add_mft(link, (np.eye(3), np.array(pt)), dt_, dr_); jt = ob.add_jt(); ob.finalize()
        set_goals(mft, jt, q)
    compare_cycles(a, b, q, dq, 4)


def test_general_hierarchies():
    # examples/06-partial_joint_task/06-partial_joint_task.cpp:108-128
    name, N = "panda_sliding_base", 16
    q, dq, _ = sample_states(name, N)
    link, pt = TASK_POINTS[name]
    S = np.zeros((2, 8)); S[0, 0] = 1; S[1, 7] = 1
    a, b = both(name, N)
    for ob in (a, b):
        ob.set_state(q, dq)
        pjt = ob.add_jt(S, name="partial_joint_task"); mft = ob.add_mft(link, (np.eye(3), np.array(pt))); ob.finalize()
        set_goals(mft, pjt, q)
    compare_cycles(a, b, q, dq, 3)
    # two partial motion-force tasks + joint task
    q, dq, _ = sample_states("panda", N)
    link, pt = TASK_POINTS["panda"]
    a, b = both("panda", N)
    for ob in (a, b):
        ob.set_state(q, dq)
        t1 = ob.add_mft(link, (np.eye(3), np.array(pt)), XYZ, [], name="position")
        t2 = ob.add_mft(link, (np.eye(3), np.array(pt)), [], XYZ, name="orientation")
        jt = ob.add_jt(); ob.finalize()
        set_goals(t1, jt, q); set_goals(t2, None, q, stream=3)
    compare_cycles(a, b, q, dq, 3)


def test_controller_checks_raise_like_the_reference():
    """RobotController.cpp:11-51: duplicate names, a task behind a full joint task"""
    for kind in ("numpy", "reference"):
        ob = OracleBatch("panda", 1, kind=kind)
        ob.add_jt(name="a"); ob.add_jt(name="b")
        with pytest.raises(ValueError):
            ob.finalize()
        ob = OracleBatch("panda", 1, kind=kind)
        ob.add_mft("end-effector", name="x"); ob.add_jt(name="x")
        with pytest.raises(ValueError):
            ob.finalize()
        ob = OracleBatch("panda", 1, kind=kind)
        jt = ob.add_jt()
        with pytest.raises(ValueError):
            jt[0].setGains(-1.0, 1.0, 0.0)            # JointTask.cpp:196-199
        with pytest.raises(ValueError):
            jt[0].setGoalPosition(np.zeros(3))         # JointTask.cpp:109-114


def _states_near_limits(robot_name, N, seed_stream=11):
    from oracle.robots import make_chain
    ch = make_chain(robot_name)
    q, dq, _ = sample_states(robot_name, N, min_sigma_ratio=0.075)
    for i in range(N):
        g = rng_for(i, stream=seed_stream)
        for _ in range(int(g.integers(0, 4))):
            j = int(g.integers(0, ch.n)); kind = int(g.integers(0, 4)); depth = g.uniform(0.005, 0.15)
            if kind == 0: q[i, j] = ch.q_upper[j] - depth
            elif kind == 1: q[i, j] = ch.q_lower[j] + depth
            elif kind == 2: dq[i, j] = ch.dq_max[j] - 3.0 * depth
            else: dq[i, j] = -ch.dq_max[j] + 3.0 * depth
    return q, dq


@pytest.mark.parametrize("hier", ["jt", "mft_jt"])
@pytest.mark.parametrize("sat,grav", [(False, False), (True, True)])
def test_joint_limit_avoidance_gravity_saturation(hier, sat, grav):
    """row f-2 + row a2 options"""
    N = 48
    q, dq = _states_near_limits("panda", N)
    link, pt = TASK_POINTS["panda"]
    a, b = both("panda", N)
    for ob in (a, b):
        ob.set_state(q, dq)
        mft = ob.add_mft(link, (np.eye(3), np.array(pt))) if hier == "mft_jt" else None
        jt = ob.add_jt(); ob.finalize()
        for c in ob.controllers:
            c.enableJointLimitAvoidance(True); c.enableGravityCompensation(grav); c.enableTorqueSaturation(sat)
        set_goals(mft, jt, q)
    compare_cycles(a, b, q, dq, 3)
    act_a = [c._joint_limit_avoidance_task._active_constraints for c in a.controllers]
    act_b = [c._joint_limit_avoidance_task._active_constraints for c in b.controllers]
    assert act_a == act_b and max(act_a) >= 2


def sensed_ex09(N, k):
    F = np.zeros((N, 3)); Mo = np.zeros((N, 3))
    for i in range(N):
        g = rng_for(i * 100003 + k, stream=33)
        F[i] = np.array([0, 0, -5.0]) + g.normal(0, 1.0, 3) * (3.0 if (k // 60) % 2 else 1.0)
        Mo[i] = g.normal(0, 0.1, 3)
    return F, Mo


def test_config3_closed_loop_force_with_passivity_400_cycles():
    """examples/09: XYZ task, force space dim 1 about Z, closed loop + POPC; the POPC window (250) and the PC period (50) are
    crossed, Rc leaves 1"""
    N, K = 3, 400
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.075, dirs=np.eye(6)[:, :3])
    link, pt = TASK_POINTS["panda"]
    a, b = both("panda", N)
    mfts = []
    for ob in (a, b):
        ob.set_state(q, dq)
        mft = ob.add_mft(link, (np.eye(3), np.array(pt)), XYZ, []); ob.add_jt(); ob.finalize()
        mfts.append(mft)
    assert rel_err(a.cycle(), b.cycle()).max() < TOL
    for mft in mfts:
        for t in mft:
            assert t.parametrizeForceMotionSpaces(1, (0, 0, 1)) is True
            t.setGoalForce((0, 0, -5.0)); t.setClosedLoopForceControl(); t.enablePassivity()
    rc_min = 1.0
    for k in range(K):
        F, Mo = sensed_ex09(N, k)
        for mft in mfts:
            for i, t in enumerate(mft):
                t.updateSensedForceAndMoment(F[i], Mo[i])
        assert rel_err(a.cycle(), b.cycle()).max() < TOL, k
        if k % 50 == 49:
            ra = [t._POPC_force._Rc for t in mfts[0]]; rb = [t._POPC_force._Rc for t in mfts[1]]
            assert np.abs(np.array(ra) - np.array(rb)).max() < 1e-12
            rc_min = min(rc_min, min(rb))
    assert rc_min < 1.0


def test_config3_surface_contact_variant():
    """examples/07: compliant-frame parametrisation, force dim 1 + moment dim 2, both loops closed, sensor frame, custom gains"""
    N, K = 6, 80
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.08)
    link, pt = TASK_POINTS["panda"]
    comp = (np.eye(3), np.array(pt))
    sensor_R = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]]); sensor_t = np.array([0.01, 0.02, 0.05])
    a, b = both("panda", N)
    mfts = []
    for ob in (a, b):
        ob.set_state(q, dq)
        mft = ob.add_mft(link, comp, in_compliant=True, name="surface_alignment_task"); ob.finalize()
        for t in mft:
            t.enablePassivity(); t.setForceSensorFrame(link, (sensor_R, sensor_t))
        set_goals(mft, None, q)
        mfts.append(mft)
    for k in range(K):
        if k == 5:
            for mft in mfts:
                for t in mft:
                    t.parametrizeForceMotionSpaces(1, (0, 0, 1)); t.parametrizeMomentRotMotionSpaces(2, (0, 0, 1))
                    t.setClosedLoopForceControl(); t.setClosedLoopMomentControl()
                    t.setGoalForce((0, 0, 10.0)); t.setGoalMoment((0, 0, 0))
                    t.setForceControlGains(0.7, 5.0, 1.5); t.setMomentControlGains(0.7, 4.0, 1.5)
        for i in range(N):
            g = rng_for(i * 7919 + k, stream=8)
            f = np.array([0.3, -0.2, 9.0]) + g.normal(0, 1.5, 3); m = g.normal(0, 0.3, 3)
            for mft in mfts:
                mft[i].updateSensedForceAndMoment(f, m)
        assert rel_err(a.cycle(), b.cycle()).max() < TOL, k
        q = q + 0.001 * dq
        a.set_state(q, dq); b.set_state(q, dq)
    for x, y in zip(*mfts):
        assert np.abs(np.array(x._sensed_force_control_world_frame) - y._sensed_force_control_world_frame).max() < 1e-12
        for name in ("sigmaForce", "sigmaPosition", "sigmaMoment", "sigmaOrientation", "getPositionError", "getOrientationError"):
            assert np.abs(np.array(getattr(x, name)()) - getattr(y, name)()).max() < 1e-12, name


def test_velocity_saturation_and_anisotropic_gains():
    N = 16
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.08)
    link, pt = TASK_POINTS["panda"]
    a, b = both("panda", N)
    for ob in (a, b):
        ob.set_state(q, dq)
        mft = ob.add_mft(link, (np.eye(3), np.array(pt))); jt = ob.add_jt(); ob.finalize()
        for x, y in zip(mft, jt):
            x.enableVelocitySaturation(0.05, 0.2); x.setPosControlGains([100, 150, 80], [20, 25, 15], [2, 0, 1]); x.setOriControlGains(150, 25, 3)
            y.enableVelocitySaturation(0.3); y.setGains(np.linspace(40, 70, 7), np.linspace(10, 16, 7), np.linspace(0, 3, 7))
        set_goals(mft, jt, q, with_vel=True)
    compare_cycles(a, b, q, dq, 3)


def test_base_transform():
    """SaiModel::TRobotBase (examples/15-17): world-frame Jacobian and pose"""
    N = 8
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.075)
    link, pt = TASK_POINTS["panda"]
    T = (rot_exp(np.array([0.3, -0.2, 0.5])), np.array([0.1, -0.4, 0.25]))
    a, b = both("panda", N, T_world_robot=T)
    for ob in (a, b):
        ob.set_state(q, dq)
        mft = ob.add_mft(link, (np.eye(3), np.array(pt))); jt = ob.add_jt(); ob.finalize()
        for c in ob.controllers:
            c.enableGravityCompensation(True)
        set_goals(mft, jt, q)
    compare_cycles(a, b, q, dq, 2)


def test_nullspace_getters():
    """TemplateTask.h:74-89"""
    N = 6
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.075)
    link, pt = TASK_POINTS["panda"]
    a, b = both("panda", N)
    tasks = []
    for ob in (a, b):
        ob.set_state(q, dq)
        mft = ob.add_mft(link, (np.eye(3), np.array(pt)), XYZ, []); jt = ob.add_jt(); ob.finalize()
        ob.cycle()
        tasks.append((mft, jt))
    for k in range(2):
        for x, y in zip(tasks[0][k], tasks[1][k]):
            for name in ("getTaskNullspace", "getPreviousTasksNullspace", "getTaskAndPreviousNullspace"):
                assert np.abs(np.array(getattr(x, name)()) - getattr(y, name)()).max() < 1e-10, name


def test_cpp_port_matches_the_reference():
    """the timed CPU baseline (oracle/cpp/osc_ref.cpp) against the reference's compiled control law, singular states included"""
    from oracle.cpp_ref import CppOracleBatch
    N = 32
    q, dq, _ = sample_states("panda", N)
    link, pt = TASK_POINTS["panda"]
    ref = OracleBatch("panda", N, kind="reference"); ref.set_state(q, dq)
    rm = ref.add_mft(link, (np.eye(3), np.array(pt))); rj = ref.add_jt(); ref.finalize()
    cb = CppOracleBatch("panda", N); cb.set_state(q, dq)
    tm = cb.add_mft(link, (np.eye(3), np.array(pt))); tj = cb.add_jt()
    x0, R0 = cb.mft_get_current(tm)
    xd = x0 + 0.03; qd = q + 0.1
    z = np.zeros((N, 3))
    cb.mft_set_goals(tm, xd, R0, z, z, z, z); cb.jt_set_goals(tj, qd)
    for i in range(N):
        rm[i].setGoalPosition(xd[i]); rj[i].setGoalPosition(qd[i])
    for c in range(3):
        assert rel_err(cb.cycle(), ref.cycle()).max() < 1e-9, c
