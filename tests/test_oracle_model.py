"""Self-consistency of the sai-model restatement (oracle/sai_model.py).  The reference pins nothing
here (no tests, sai-model absent), so the restatement is checked against first principles:
finite differences, the kinetic-energy identity, closed forms for the planar arm."""
import numpy as np
import pytest

from oracle.robots import make_chain
from oracle.sai_model import SaiModel, computePseudoInverse, matrixRangeBasis, orientationError
from tests.osc_testlib import TASK_POINTS, rng_for

ROBOTS = ["panda", "rrrr", "puma_like", "panda_sliding_base"]


def _state(name, i=0):
    ch = make_chain(name)
    g = rng_for(i, stream=11)
    q = ch.q_lower + (0.1 + 0.8 * g.random(ch.n)) * (ch.q_upper - ch.q_lower)
    dq = g.uniform(-1, 1, ch.n)
    m = SaiModel(ch)
    m.setQ(q); m.setDq(dq); m.updateModel()
    return ch, m, q, dq


@pytest.mark.parametrize("name", ROBOTS)
def test_jacobian_matches_finite_differences(name):
    ch, m, q, dq = _state(name)
    link, pt = TASK_POINTS[name]
    J = m.J(link, pt)
    eps = 1e-6
    for i in range(ch.n):
        qp, qm = q.copy(), q.copy()
        qp[i] += eps; qm[i] -= eps
        m.setQ(qp); m.updateKinematics(); xp, Rp = m.position(link, pt), m.rotation(link)
        m.setQ(qm); m.updateKinematics(); xm, Rm = m.position(link, pt), m.rotation(link)
        assert np.abs((xp - xm) / (2 * eps) - J[:3, i]).max() < 1e-8
        W = (Rp - Rm) / (2 * eps) @ m.rotation(link).T   # ~ skew(omega) up to O(eps^2)
        m.setQ(q); m.updateKinematics()
        W = (Rp - Rm) / (2 * eps) @ m.rotation(link).T
        w = np.array([W[2, 1], W[0, 2], W[1, 0]])
        assert np.abs(w - J[3:, i]).max() < 1e-7


@pytest.mark.parametrize("name", ROBOTS)
def test_mass_matrix_energy_identity_and_spd(name):
    for i in range(5):
        ch, m, q, dq = _state(name, i)
        M = m.M()
        assert np.abs(M - M.T).max() == 0
        assert np.linalg.eigvalsh(M).min() > 0
        assert abs(0.5 * dq @ M @ dq - m.kineticEnergy()) < 1e-12 * max(1.0, m.kineticEnergy())
        assert np.abs(M @ m.MInv() - np.eye(ch.n)).max() < 1e-10


@pytest.mark.parametrize("name", ROBOTS)
def test_gravity_is_gradient_of_potential(name):
    ch, m, q, dq = _state(name)
    g = m.jointGravityVector()

    def pot(qq):
        m.setQ(qq); m.updateKinematics()
        return sum(ch.mass[k] * 9.81 * (m._pb[k] + m._Rb[k] @ ch.com[k])[2] for k in range(ch.n))
    eps = 1e-6
    fd = np.array([(pot(q + eps * np.eye(ch.n)[i]) - pot(q - eps * np.eye(ch.n)[i])) / (2 * eps) for i in range(ch.n)])
    assert np.abs(fd - g).max() < 1e-6


def test_planar_arm_closed_forms():
    """RRRR at q = 0 is a straight 2 m bar along x: closed-form Jacobian and M[3,3]."""
    ch = make_chain("rrrr")
    m = SaiModel(ch)
    J = m.J("link4", (0.5, 0, 0))
    assert np.allclose(J[1], [2.0, 1.5, 1.0, 0.5], atol=1e-15)      # d y / d q_i = distance to the tip
    assert np.allclose(J[0], 0) and np.allclose(J[5], 1)
    M = m.M()
    assert abs(M[3, 3] - (0.000967 + 0.25 ** 2)) < 1e-15              # izz + m c^2
    # joint 1 carries 4 links: sum over links of izz + m d^2
    d = [0.25, 0.75, 1.25, 1.75]
    assert abs(M[0, 0] - sum(0.000967 + x * x for x in d)) < 1e-13


def test_panda_facts_from_survey():
    """M[6,6] = 0.07 + 0.01 = 0.08 < 0.1 in every configuration (BIE always clamps joint 7)."""
    for i in range(5):
        ch, m, q, dq = _state("panda", i)
        assert abs(m.M()[6, 6] - 0.08) < 1e-14
    assert abs(ch.mass[6] - 2.0) < 1e-15   # link7 1.8 + end-effector 0.2 merged (RBDL fixed-joint merge)


@pytest.mark.parametrize("name", ROBOTS)
def test_operational_space_identities(name):
    ch, m, q, dq = _state(name, 3)
    link, pt = TASK_POINTS[name]
    J = m.J(link, pt)
    if name == "rrrr":
        J = J[[0, 1, 5]]
    o = m.operationalSpaceMatrices(J)
    n = ch.n
    assert np.abs(J @ o.N).max() < 1e-9                               # J N = 0
    assert np.abs(o.N @ o.N - o.N).max() < 1e-9                       # N idempotent
    assert np.abs(J @ o.Jbar - np.eye(J.shape[0])).max() < 1e-9       # J Jbar = I
    assert np.abs(o.Lambda - o.Lambda.T).max() < 1e-9 * np.abs(o.Lambda).max()


def test_free_functions():
    A = np.diag([3.0, 2.0, 1e-5])
    U = matrixRangeBasis(A)
    assert U.shape == (3, 2)
    assert matrixRangeBasis(np.eye(3)).shape == (3, 3) and np.array_equal(matrixRangeBasis(np.eye(3)), np.eye(3))
    assert np.linalg.norm(matrixRangeBasis(np.zeros((3, 3)))) == 0
    R = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    e = orientationError(R, np.eye(3))        # desired = 90 deg about z
    assert np.allclose(e, [0, 0, -1.0])        # -1/2 sum Rc_i x Rd_i
    with pytest.raises(ValueError):
        orientationError(2 * np.eye(3), np.eye(3))
    K = np.diag([20.0, 0.0, 5.0])
    assert np.allclose(computePseudoInverse(K), np.diag([0.05, 0.0, 0.2]))


def test_joint_limit_avoidance_oracle_properties():
    """JointLimitAvoidanceTask restatement (reference JointLimitAvoidanceTask.cpp:124-421): status zones, the constraint
    null space and the RobotController blend (RobotController.cpp:96-112)."""
    from oracle import primitives as OP
    from oracle.robots import make_chain
    from oracle.sai_model import SaiModel
    ch = make_chain("panda")
    m = SaiModel(ch)
    q = ch.q_lower + 0.5 * (ch.q_upper - ch.q_lower)
    dq = np.zeros(ch.n)
    q[3] = ch.q_upper[3] - 0.05          # inside zone 2 (6 deg = 0.1047)
    q[1] = ch.q_lower[1] + 0.13          # inside zone 1 (9 deg = 0.157) only
    dq[5] = ch.dq_max[5] - 0.2           # velocity zone 2 (0.3)
    m.setQ(q); m.setDq(dq); m.updateModel()
    jt = OP.JointTask(m)
    c = OP.RobotController(m, [jt])
    jt.setGoalPosition(q + 0.3)
    c.updateControllerTaskModels()
    plain = c.computeControlTorques()
    c.enableJointLimitAvoidance(True)
    c.updateControllerTaskModels()
    tau = c.computeControlTorques()
    jla = c._joint_limit_avoidance_task
    assert jla._limit_status == [jla.OFF, jla.POS_Z1, jla.OFF, jla.POS_Z2, jla.OFF, jla.VEL_Z2, jla.OFF]
    assert jla._limit_direction[1] == jla.NEGATIVE and jla._limit_direction[3] == jla.POSITIVE
    S = jla._joint_selection
    N = c._N_constraints
    assert np.allclose(N @ N, N, atol=1e-12)                    # projector
    assert np.allclose(S @ N, 0.0, atol=1e-12)                  # no acceleration of the constrained joints from the tasks
    inactive = [0, 2, 4, 6]
    assert np.allclose(tau[inactive], plain[inactive], atol=1e-12)   # N^T only touches the active joints' torques
    assert tau[3] < plain[3]                                    # pushed away from the upper limit
    # far from every limit the task does nothing
    m.setQ(ch.q_lower + 0.5 * (ch.q_upper - ch.q_lower)); m.setDq(np.zeros(ch.n)); m.updateModel()
    c.updateControllerTaskModels()
    a = c.computeControlTorques()
    c.enableJointLimitAvoidance(False)
    assert np.array_equal(a, c.computeControlTorques())


@pytest.mark.parametrize("robot_name", ["panda", "panda_sliding_base", "rrrr"])
def test_bias_forces_against_independent_statements(robot_name):
    """oracle/simulation.py: recursive Newton-Euler bias forces vs (a) the gravity vector from the
    Jacobians, (b) Coriolis/centrifugal forces from the Christoffel symbols of a finite-difference dM/dq, (c) the power
    balance d/dt(1/2 dq^T M dq) = dq^T (tau - g) along an integrated trajectory."""
    from oracle import simulation as SIM
    from oracle.robots import make_chain
    from oracle.sai_model import SaiModel
    ch = make_chain(robot_name)
    rng = np.random.default_rng(7)
    m = SaiModel(ch)
    q = ch.q_lower + rng.random(ch.n) * (ch.q_upper - ch.q_lower)
    dq = rng.uniform(-1, 1, ch.n)
    m.setQ(q); m.setDq(np.zeros(ch.n)); m.updateModel()
    assert np.allclose(SIM.bias_forces(m), m.jointGravityVector(), atol=1e-11)
    m.setDq(dq); m.updateModel()
    b = SIM.bias_forces(m) - m.jointGravityVector()
    h = 1e-6
    dM = np.zeros((ch.n, ch.n, ch.n))
    for k in range(ch.n):
        e = np.zeros(ch.n); e[k] = h
        m.setQ(q + e); m.updateModel(); Mp = m.M()
        m.setQ(q - e); m.updateModel(); Mm = m.M()
        dM[:, :, k] = (Mp - Mm) / (2 * h)
    m.setQ(q); m.updateModel()
    c_ref = np.einsum("ijk,j,k->i", dM, dq, dq) - 0.5 * np.einsum("jki,j,k->i", dM, dq, dq)
    assert np.abs(b - c_ref).max() < 1e-6 * max(1.0, np.abs(c_ref).max())
    # power balance over a short trajectory with a constant torque
    tau = rng.uniform(-2, 2, ch.n)
    e0 = 0.5 * m.dq() @ m.M() @ m.dq()
    work = 0.0
    dt = 1e-5
    for _ in range(200):
        dq0 = m.dq(); g = m.jointGravityVector()
        SIM.integrate(m, tau, dt)
        work += 0.5 * (dq0 + m.dq()) @ (tau - g) * dt
    e1 = 0.5 * m.dq() @ m.M() @ m.dq()
    assert abs((e1 - e0) - work) < 2e-4 * max(1.0, abs(work))
