"""Generates tests/golden/config*.npz FROM THE REFERENCE'S OWN COMPILED CONTROL LAW.

Source: oracle/_ref/libsai_ref_orient.so = /root/reference/src/{RobotController,tasks/JointTask,tasks/MotionForceTask,
tasks/SingularityHandler,tasks/JointLimitAvoidanceTask,helper_modules/POPCExplicitForceControl,...}.cpp compiled where they
lie, unmodified (oracle/Makefile, oracle/sai_ref.py); the model arithmetic under it is the sai-model stand-in.  The reference
ships no golden vectors of its own, so these are "outputs of the reference itself run here".  Each file records
`source`.  Where the singular branch is exercised the file also holds `tau_eigen_signs`: the same run with the stand-in's
JacobiSVD left at the signs Eigen's published algorithm produces (libsai_ref.so) instead of this repository's orientation
convention, and `sign_sensitive`: the robots whose torques differ between the two (classifySingularity perturbs q along
+V_s, SingularityHandler.cpp:254, so type-1/type-2 depends on the sign of the singular vector).

Regenerate with:  python tests/golden/generate.py      (needs /root/reference; `--source numpy` freezes the restatement instead)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import osc_testlib  # noqa: E402
from tests.osc_testlib import TASK_POINTS, rng_for, rot_exp, sample_states, sensed_ex09  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SOURCE = "reference"   # "reference": oracle/_ref/libsai_ref_orient.so; "numpy": oracle/primitives.py


def OracleBatch(name, N, eigen_signs=False):
    if SOURCE == "reference" and eigen_signs:
        from oracle.sai_ref import RefBatch
        return RefBatch(name, N, oriented=False)
    return osc_testlib.OracleBatch(name, N, kind=SOURCE)


def source_tag():
    return np.array("reference: /root/reference/src compiled in place (oracle/_ref/libsai_ref_orient.so)" if SOURCE == "reference"
                    else "numpy restatement (oracle/primitives.py)")


def goals_for(N, n, x0, R0, q, stream):
    xd = np.zeros((N, 3)); Rd = np.zeros((N, 3, 3)); vd = np.zeros((N, 3)); wd = np.zeros((N, 3)); ad = np.zeros((N, 3)); ald = np.zeros((N, 3))
    qd = np.zeros((N, n))
    for i in range(N):
        g = rng_for(i, stream=stream)
        xd[i] = x0[i] + g.uniform(-0.05, 0.05, 3); Rd[i] = R0[i] @ rot_exp(g.uniform(-0.2, 0.2, 3))
        vd[i] = g.uniform(-0.1, 0.1, 3); wd[i] = g.uniform(-0.1, 0.1, 3); ad[i] = g.uniform(-0.5, 0.5, 3); ald[i] = g.uniform(-0.5, 0.5, 3)
        qd[i] = q[i] + g.uniform(-0.2, 0.2, n)
    return dict(xd=xd, Rd=Rd, vd=vd, wd=wd, ad=ad, ald=ald, qd=qd)


def apply_goals(omft, ojt, G):
    for i, t in enumerate(omft):
        t.setGoalPosition(G["xd"][i]); t.setGoalOrientation(G["Rd"][i]); t.setGoalLinearVelocity(G["vd"][i])
        t.setGoalAngularVelocity(G["wd"][i]); t.setGoalLinearAcceleration(G["ad"][i]); t.setGoalAngularAcceleration(G["ald"][i])
    for i, t in enumerate(ojt):
        t.setGoalPosition(G["qd"][i])


def config1():
    """Panda, single JointTask, kp 100 kv 20 ki 3, BIE; 3 cycles"""
    N = 16
    q, dq, _ = sample_states("panda", N)
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    ojt = ob.add_jt(); ob.finalize()
    qd = np.zeros((N, 7))
    for i, t in enumerate(ojt):
        t.setGains(100.0, 20.0, 3.0)
        qd[i] = q[i] + rng_for(i, stream=31).uniform(-0.2, 0.2, 7)
        t.setGoalPosition(qd[i])
    tau = np.array([ob.cycle() for _ in range(3)])
    np.savez(os.path.join(OUT, "config1_joint_task.npz"), q=q, dq=dq, qd=qd, tau=tau, source=source_tag())


def _osc_run(name, N, q, dq, dt_, dr_, stream, eigen_signs=False):
    link, pt = TASK_POINTS[name]
    ob = OracleBatch(name, N, eigen_signs=eigen_signs); ob.set_state(q, dq)
    omft = ob.add_mft(link, (np.eye(3), np.array(pt)), dt_, dr_); ojt = ob.add_jt(); ob.finalize()
    x0 = np.array([t._current_position for t in omft]); R0 = np.array([t._current_orientation for t in omft])
    G = goals_for(N, q.shape[1], x0, R0, q, stream)
    apply_goals(omft, ojt, G)
    tau = np.array([ob.cycle() for _ in range(3)])
    singular = np.array([len(t._singularity_handler._singularity_types) != 0 for t in omft])
    s = [np.asarray(t._singularity_handler._svd_s) for t in omft]
    smin = np.array([x[-1] / x[0] for x in s])
    return tau, singular, smin, x0, R0, G


def _sign_sensitive(tau, tau_e):
    return (np.abs(tau - tau_e).max(axis=(0, 2)) > 1e-9 * np.maximum(np.abs(tau).max(axis=(0, 2)), 1e-9))


def config2():
    """Panda, MotionForceTask 6-DoF + JointTask null space via RobotController, defaults (BIE); includes singular states"""
    N = 32
    q, dq, _ = sample_states("panda", N)     # unfiltered: ~half of the robots take the blending branch
    tau, singular, smin, x0, R0, G = _osc_run("panda", N, q, dq, None, None, 32)
    extra = {}
    if SOURCE == "reference":
        tau_e = _osc_run("panda", N, q, dq, None, None, 32, eigen_signs=True)[0]
        extra = dict(tau_eigen_signs=tau_e, sign_sensitive=_sign_sensitive(tau, tau_e))
    np.savez(os.path.join(OUT, "config2_osc_nullspace.npz"), q=q, dq=dq, tau=tau, singular=singular, sigma_ratio=smin, x0=x0, R0=R0,
             source=source_tag(), **extra, **G)


def config3():
    """Panda, XYZ task, force space dim 1 about Z, closed loop + passivity (ex.09), JointTask in the null space,
    320 cycles with a noisy sensed force so that the POPC window (250) and PC period (50) are exercised"""
    N = 4
    K = 320
    dirs = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.075, dirs=np.eye(6)[:, :3])
    link, pt = TASK_POINTS["panda"]
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    omft = ob.add_mft(link, (np.eye(3), np.array(pt)), dirs, []); ojt = ob.add_jt(); ob.finalize()
    tau0 = ob.cycle()
    for t in omft:
        t.parametrizeForceMotionSpaces(1, (0, 0, 1)); t.setGoalForce((0, 0, -5.0)); t.setClosedLoopForceControl(); t.enablePassivity()
    F = np.zeros((K, N, 3)); Mo = np.zeros((K, N, 3)); tau = np.zeros((K, N, 7)); rc = np.zeros((K, N))
    for k in range(K):
        F[k], Mo[k] = sensed_ex09(N, k)
        for i in range(N):
            omft[i].updateSensedForceAndMoment(F[k, i], Mo[k, i])
        tau[k] = ob.cycle()
        rc[k] = [t._POPC_force._Rc for t in omft]
    np.savez(os.path.join(OUT, "config3_force_popc.npz"), q=q, dq=dq, tau0=tau0, sensed_force=F, sensed_moment=Mo, tau=tau, Rc=rc, source=source_tag())


CONFIG3_CHECK_CYCLES = (1, 50, 51, 250, 251, 300, 1000)


def config3_1000():
    """SURVEY.md 8(d) config 3 as written: the ex.09 set-up run for K = 1000 consecutive cycles (sensed wrench regenerated from
    the seed by osc_testlib.sensed_ex09), torques kept at cycles {1, 50, 51, 250, 251, 300, 1000} plus every 100th, Rc at every
    PC update; the state moves a little every cycle so that the model stage is re-evaluated"""
    N, K = 8, 1000
    dirs = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
    q0, dq, _ = sample_states("panda", N, min_sigma_ratio=0.1, dirs=np.eye(6)[:, :3])
    link, pt = TASK_POINTS["panda"]
    ob = OracleBatch("panda", N); ob.set_state(q0, dq)
    omft = ob.add_mft(link, (np.eye(3), np.array(pt)), dirs, []); ojt = ob.add_jt(); ob.finalize()
    for t in omft:
        t.parametrizeForceMotionSpaces(1, (0, 0, 1)); t.setGoalForce((0, 0, -5.0)); t.setClosedLoopForceControl(); t.enablePassivity()
    keep = sorted(set(CONFIG3_CHECK_CYCLES) | set(range(100, K + 1, 100)))
    tau = np.zeros((len(keep), N, 7)); rc = np.zeros((K // 50, N)); popc = None
    q = q0.copy()
    for k in range(1, K + 1):
        F, Mo = sensed_ex09(N, k - 1)
        for i in range(N):
            omft[i].updateSensedForceAndMoment(F[i], Mo[i])
        t = ob.cycle()
        if k in keep:
            tau[keep.index(k)] = t
        if k % 50 == 0:
            rc[k // 50 - 1] = [x._POPC_force._Rc for x in omft]
        q = q0 + 0.05 * np.sin(2 * np.pi * k / 500.0) * dq      # bounded excursion around the sampled state
        ob.set_state(q, dq)
    popc = np.array([[x._POPC_force._passivity_observer_value, x._POPC_force._E_correction, x._POPC_force._Rc] for x in omft])
    np.savez(os.path.join(OUT, "config3_force_popc_1000.npz"), q=q0, dq=dq, cycles=np.array(keep), tau=tau, Rc=rc, popc_final=popc, source=source_tag())


def config4():
    """mixed-DoF: RRRR planar task (ex.11) and PUMA-like full task, unfiltered states, 3 cycles"""
    out = {}
    for name, dt_, dr_ in (("rrrr", [(1, 0, 0), (0, 1, 0)], [(0, 0, 1)]), ("puma_like", None, None)):
        N = 16
        q, dq, _ = sample_states(name, N)
        tau, singular, smin, x0, R0, G = _osc_run(name, N, q, dq, dt_, dr_, 34)
        extra = {}
        if SOURCE == "reference":
            tau_e = _osc_run(name, N, q, dq, dt_, dr_, 34, eigen_signs=True)[0]
            extra = dict(tau_eigen_signs=tau_e, sign_sensitive=_sign_sensitive(tau, tau_e))
        out.update({name + "_" + k: v for k, v in dict(q=q, dq=dq, tau=tau, singular=singular, **extra, **G).items()})
    np.savez(os.path.join(OUT, "config4_mixed_dof.npz"), source=source_tag(), **out)


def singular_replay_states(N=12):
    """Panda states inside the blending band whose motion keeps them there for hundreds of cycles"""
    q, dq, _ = sample_states("panda", 200)
    from oracle.sai_model import SaiModel
    from oracle.robots import make_chain
    model = SaiModel(make_chain("panda"))
    link, pt = TASK_POINTS["panda"]
    pick = []
    for i in range(200):
        ok = True
        for k in (0, 150, 300):
            model.setQ(q[i] + 0.0002 * k * dq[i]); model.updateKinematics()
            s = np.linalg.svd(model.J(link, pt), compute_uv=False)
            ok = ok and (8e-3 < s[5] / s[0] < 5e-2)
        if ok:
            pick.append(i)
        if len(pick) == N:
            break
    return q[pick], dq[pick]


def config4_singular_replay():
    """SingularityHandler.cpp:276-293 driven past its 200-entry history: 300 consecutive cycles inside the blending band, the state
    drifting every cycle; torques of every cycle, the type counters and the history length at cycles 100/200/201/260/300"""
    K = 300
    q0, dq = singular_replay_states()
    N = q0.shape[0]
    link, pt = TASK_POINTS["panda"]
    res = {}
    for tag, eigen_signs in (("", False), ("_eigen_signs", True)):
        if eigen_signs and SOURCE != "reference":
            continue
        ob = OracleBatch("panda", N, eigen_signs=eigen_signs); ob.set_state(q0, dq)
        omft = ob.add_mft(link, (np.eye(3), np.array(pt))); ojt = ob.add_jt(); ob.finalize()
        x0 = np.array([t._current_position for t in omft]); R0 = np.array([t._current_orientation for t in omft])
        G = goals_for(N, 7, x0, R0, q0, 35)
        apply_goals(omft, ojt, G)
        tau = np.zeros((K, N, 7)); counters = {}
        for k in range(1, K + 1):
            tau[k - 1] = ob.cycle()
            if k in (100, 200, 201, 260, 300):
                h = [t._singularity_handler for t in omft]
                counters[k] = np.array([[x._type_1_counter, x._type_2_counter] for x in h])
            ob.set_state(q0 + 0.0002 * k * dq, dq)
        res["tau" + tag] = tau
        res["counters" + tag] = np.array([counters[k] for k in (100, 200, 201, 260, 300)])
        if not eigen_signs:
            res.update(G)
    np.savez_compressed(os.path.join(OUT, "config4_singular_replay.npz"), q=q0, dq=dq, counter_cycles=np.array([100, 200, 201, 260, 300]),
                        source=source_tag(), **res)


if __name__ == "__main__":
    if "--source" in sys.argv:
        SOURCE = sys.argv[sys.argv.index("--source") + 1]
    config1(); config2(); config3(); config3_1000(); config4(); config4_singular_replay()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
