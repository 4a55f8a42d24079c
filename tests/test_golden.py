"""Golden fixtures tests/golden/config*.npz: outputs of THE REFERENCE'S OWN COMPILED CONTROL LAW (oracle/_ref/libsai_ref_orient.so,
the reference's sources compiled where they lie -- see tests/golden/generate.py; the reference itself ships no vectors).
CPU: the numpy restatement and the C++ port reproduce them, and the files are bit-identical to what the library produces where
it exists.  GPU: the CUDA path reproduces them through the C ABI without touching oracle/ at run time."""
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


def test_fixtures_come_from_the_reference_and_regenerate_bit_identically(tmp_path):
    for f in ("config1_joint_task", "config2_osc_nullspace", "config3_force_popc", "config3_force_popc_1000", "config4_mixed_dof", "config4_singular_replay"):
        assert str(load(f + ".npz")["source"]).startswith("reference: /root/reference/src compiled in place")
    from oracle import sai_ref
    if not (sai_ref.available(True) and sai_ref.available(False)):
        pytest.skip("oracle/_ref/libsai_ref*.so not built here")
    from tests.golden import generate
    old = generate.OUT
    generate.OUT = str(tmp_path)
    try:
        generate.config1(); generate.config2(); generate.config4()
    finally:
        generate.OUT = old
    for f in ("config1_joint_task.npz", "config2_osc_nullspace.npz", "config4_mixed_dof.npz"):
        a, b = load(f), np.load(os.path.join(str(tmp_path), f))
        for k in a.files:
            assert np.array_equal(a[k], b[k]), (f, k)


def test_svd_sign_sensitivity_of_the_fixtures():
    """classifySingularity perturbs q along +V_s (SingularityHandler.cpp:254): with the signs Eigen's JacobiSVD procedure yields
    (tau_eigen_signs) instead of this repository's orientation convention (tau), only robots on the singular branch can differ,
    and only some of them do (DESIGN.md section 3 lists the counts)."""
    d = load("config2_osc_nullspace.npz")
    assert d["sign_sensitive"].sum() == 4 and d["singular"].sum() == 16
    assert not d["sign_sensitive"][~d["singular"]].any()
    same = ~d["sign_sensitive"]
    assert rel_err(d["tau_eigen_signs"][0][same], d["tau"][0][same]).max() < 1e-9
    m = load("config4_mixed_dof.npz")
    assert m["puma_like_sign_sensitive"].sum() == 1 and m["rrrr_sign_sensitive"].sum() == 0
    r = load("config4_singular_replay.npz")
    differs = np.abs(r["tau"] - r["tau_eigen_signs"]).max(axis=(0, 2)) > 1e-9
    assert differs.sum() == 2
    assert np.array_equal(r["counters"].sum(axis=2), r["counters_eigen_signs"].sum(axis=2))


def _numpy_osc(name, q, dq, dt_, dr_, goals):
    from tests.osc_testlib import OracleBatch
    N = q.shape[0]
    link, pt = TASK_POINTS[name]
    ob = OracleBatch(name, N, kind="numpy"); ob.set_state(q, dq)
    omft = ob.add_mft(link, (np.eye(3), np.array(pt)), dt_, dr_); ojt = ob.add_jt(); ob.finalize()
    if goals is not None:
        for i in range(N):
            t = omft[i]
            t.setGoalPosition(goals["xd"][i]); t.setGoalOrientation(goals["Rd"][i]); t.setGoalLinearVelocity(goals["vd"][i])
            t.setGoalAngularVelocity(goals["wd"][i]); t.setGoalLinearAcceleration(goals["ad"][i]); t.setGoalAngularAcceleration(goals["ald"][i])
            ojt[i].setGoalPosition(goals["qd"][i])
    return ob, omft, ojt


def test_numpy_restatement_reproduces_the_reference_fixtures():
    d = load("config2_osc_nullspace.npz")
    ob, _, _ = _numpy_osc("panda", d["q"], d["dq"], None, None, d)
    for k in range(3):
        assert rel_err(ob.cycle(), d["tau"][k]).max() < 1e-10
    m = load("config4_mixed_dof.npz")
    for name, dt_, dr_ in (("rrrr", [(1, 0, 0), (0, 1, 0)], [(0, 0, 1)]), ("puma_like", None, None)):
        goals = {k: m[name + "_" + k] for k in ("xd", "Rd", "vd", "wd", "ad", "ald", "qd")}
        ob, _, _ = _numpy_osc(name, m[name + "_q"], m[name + "_dq"], dt_, dr_, goals)
        for k in range(3):
            assert rel_err(ob.cycle(), m[name + "_tau"][k]).max() < 1e-10
    # history eviction: the first 3 robots of the 300-cycle singular replay
    r = load("config4_singular_replay.npz")
    sel = slice(0, 3)
    goals = {k: r[k][sel] for k in ("xd", "Rd", "vd", "wd", "ad", "ald", "qd")}
    ob, omft, _ = _numpy_osc("panda", r["q"][sel], r["dq"][sel], None, None, goals)
    for k in range(1, 301):
        assert rel_err(ob.cycle(), r["tau"][k - 1][sel]).max() < 1e-10, k
        if k in r["counter_cycles"]:
            c = r["counters"][list(r["counter_cycles"]).index(k)][sel]
            assert [[t._singularity_handler._type_1_counter, t._singularity_handler._type_2_counter] for t in omft] == c.tolist()
        ob.set_state(r["q"][sel] + 0.0002 * k * r["dq"][sel], r["dq"][sel])
    assert len(omft[0]._singularity_handler._singularity_history) == 200      # SingularityHandler.cpp:286-293


def _replay_config3_1000(cycle_fn, set_state_fn, sensed_fn, d, robots=None):
    """drives the 1000-cycle ex.09 scenario of tests/golden/generate.py::config3_1000 and yields (k, tau)"""
    from tests.osc_testlib import sensed_ex09
    q0, dq = d["q"], d["dq"]
    N = q0.shape[0]
    for k in range(1, 1001):
        F, Mo = sensed_ex09(N, k - 1)
        sensed_fn(F, Mo)
        yield k, cycle_fn()
        set_state_fn(q0 + 0.05 * np.sin(2 * np.pi * k / 500.0) * dq, dq)


def test_cpp_port_reproduces_config3_1000_cycles():
    """SURVEY.md 8(d) config 3 as written: K = 1000 cycles, torques at {1, 50, 51, 250, 251, 300, 1000}"""
    from oracle.cpp_ref import CppOracleBatch
    d = load("config3_force_popc_1000.npz")
    N = d["q"].shape[0]
    link, pt = TASK_POINTS["panda"]
    cb = CppOracleBatch("panda", N); cb.set_state(d["q"], d["dq"])
    tm = cb.add_mft(link, (np.eye(3), np.array(pt)), [(1, 0, 0), (0, 1, 0), (0, 0, 1)], []); cb.add_jt()
    cb.mft_force_setup(tm, fdim=1, faxis=(0, 0, 1), cl_force=True, passivity=True)
    cb.mft_set_force_goals(tm, np.tile([0, 0, -5.0], (N, 1)), np.zeros((N, 3)))
    keep = list(d["cycles"])
    assert set((1, 50, 51, 250, 251, 300, 1000)) <= set(keep)
    for k, tau in _replay_config3_1000(cb.cycle, cb.set_state, lambda F, Mo: cb.mft_update_sensed(tm, F, Mo), d):
        if k in keep:
            assert rel_err(tau, d["tau"][keep.index(k)]).max() < 1e-9, k
    assert d["Rc"].min() == 0.0 and d["Rc"][-1].max() > 0.4       # the passivity controller saturated and relaxed again


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
        same = ~d["sign_sensitive"]                           # the Eigen-procedure signs agree wherever the sign does not matter
        assert rel_err(tau[same], d["tau_eigen_signs"][k][same]).max() < REL_TOL
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


@pytest.mark.gpu
def test_gpu_reproduces_config3_1000_cycles():
    """SURVEY.md 8(d) config 3 as written: the ex.09 closed-loop force + POPC scenario for K = 1000 consecutive cycles with a
    moving state, torques checked at cycles {1, 50, 51, 250, 251, 300, 1000} (and every 100th) against the reference's output,
    Rc at every passivity-controller update"""
    import sai_primitives_b200 as sp
    d = load("config3_force_popc_1000.npz")
    robot, mft, jt, ctrl = _gpu_osc(sp, "panda", d["q"], d["dq"], [(1, 0, 0), (0, 1, 0), (0, 0, 1)], [], None)
    assert mft.parametrizeForceMotionSpaces(1, (0, 0, 1)) is True
    mft.setGoalForce(np.array([0, 0, -5.0])); mft.setClosedLoopForceControl(); mft.enablePassivity()
    keep = list(d["cycles"])

    def cycle():
        ctrl.updateControllerTaskModels()
        return ctrl.computeControlTorques()

    def set_state(q, dq):
        robot.setQ(q); robot.setDq(dq); robot.updateModel()

    checked = 0
    for k, tau in _replay_config3_1000(cycle, set_state, mft.updateSensedForceAndMoment, d):
        if k in keep:
            assert rel_err(tau, d["tau"][keep.index(k)]).max() < REL_TOL, k
            checked += 1
        if k % 50 == 0:
            assert np.abs(mft._get(sp.capi.MFT_POPC_STATE)[:, 2] - d["Rc"][k // 50 - 1]).max() < 1e-9, k
    assert checked == len(keep)
    popc = mft._get(sp.capi.MFT_POPC_STATE)
    assert np.abs(popc[:, 0] - d["popc_final"][:, 0]).max() < 1e-9 * np.abs(d["popc_final"][:, 0]).max()
    assert (robot.status() & (sp.capi.STATUS_POPC_OVERFLOW | sp.capi.STATUS_UNHANDLED)).sum() == 0


@pytest.mark.gpu
def test_gpu_reproduces_singular_replay_with_history_eviction():
    """SingularityHandler.cpp:276-293: 300 consecutive cycles inside the blending band, so that the 200-entry classification
    history evicts; every cycle's torques against the reference's output"""
    import sai_primitives_b200 as sp
    r = load("config4_singular_replay.npz")
    robot, mft, jt, ctrl = _gpu_osc(sp, "panda", r["q"], r["dq"], None, None, r)
    worst = 0.0
    for k in range(1, 301):
        ctrl.updateControllerTaskModels()
        tau = ctrl.computeControlTorques()
        st = robot.status()
        assert (st & sp.capi.STATUS_UNHANDLED).sum() == 0 and ((st & sp.capi.STATUS_SINGULAR_PATH) != 0).all()
        worst = max(worst, rel_err(tau, r["tau"][k - 1]).max())
        assert worst < REL_TOL, k
        robot.setQ(r["q"] + 0.0002 * k * r["dq"]); robot.updateModel()
