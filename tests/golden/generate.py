"""Generates tests/golden/*.npz from the numpy restatement (oracle/primitives.py).

PARITY UNPINNED: the reference ships no golden vectors and cannot be run here, so
these fixtures freeze the ORACLE's outputs (not the reference's) on seeded inputs.  They pin the
oracle against accidental change and give the CUDA path a check that does not need the oracle at
run time.  Regenerate with:  python tests/golden/generate.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.osc_testlib import TASK_POINTS, OracleBatch, rng_for, rot_exp, sample_states  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


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
    np.savez(os.path.join(OUT, "config1_joint_task.npz"), q=q, dq=dq, qd=qd, tau=tau)


def config2():
    """Panda, MotionForceTask 6-DoF + JointTask null space via RobotController, defaults (BIE); includes singular states"""
    N = 32
    q, dq, _ = sample_states("panda", N)     # unfiltered: ~half of the robots take the blending branch
    link, pt = TASK_POINTS["panda"]
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    omft = ob.add_mft(link, (np.eye(3), np.array(pt))); ojt = ob.add_jt(); ob.finalize()
    x0 = np.array([t._current_position for t in omft]); R0 = np.array([t._current_orientation for t in omft])
    G = goals_for(N, 7, x0, R0, q, 32)
    apply_goals(omft, ojt, G)
    tau = np.array([ob.cycle() for _ in range(3)])
    singular = np.array([len(t._singularity_handler._singularity_types) != 0 for t in omft])
    smin = np.array([t._singularity_handler._svd_s[5] / t._singularity_handler._svd_s[0] for t in omft])
    np.savez(os.path.join(OUT, "config2_osc_nullspace.npz"), q=q, dq=dq, tau=tau, singular=singular, sigma_ratio=smin, x0=x0, R0=R0, **G)


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
        for i in range(N):
            g = rng_for(i * 100003 + k, stream=33)
            F[k, i] = np.array([0, 0, -5.0]) + g.normal(0, 1.0, 3) * (3.0 if (k // 60) % 2 else 1.0)
            Mo[k, i] = g.normal(0, 0.1, 3)
            omft[i].updateSensedForceAndMoment(F[k, i], Mo[k, i])
        tau[k] = ob.cycle()
        rc[k] = [t._POPC_force._Rc for t in omft]
    np.savez(os.path.join(OUT, "config3_force_popc.npz"), q=q, dq=dq, tau0=tau0, sensed_force=F, sensed_moment=Mo, tau=tau, Rc=rc)


def config4():
    """mixed-DoF: RRRR planar task (ex.11) and PUMA-like full task, unfiltered states, 3 cycles"""
    out = {}
    for name, dt_, dr_ in (("rrrr", [(1, 0, 0), (0, 1, 0)], [(0, 0, 1)]), ("puma_like", None, None)):
        N = 16
        q, dq, _ = sample_states(name, N)
        link, pt = TASK_POINTS[name]
        ob = OracleBatch(name, N); ob.set_state(q, dq)
        omft = ob.add_mft(link, (np.eye(3), np.array(pt)), dt_, dr_); ojt = ob.add_jt(); ob.finalize()
        x0 = np.array([t._current_position for t in omft]); R0 = np.array([t._current_orientation for t in omft])
        G = goals_for(N, q.shape[1], x0, R0, q, 34)
        apply_goals(omft, ojt, G)
        tau = np.array([ob.cycle() for _ in range(3)])
        singular = np.array([len(t._singularity_handler._singularity_types) != 0 for t in omft])
        out.update({name + "_" + k: v for k, v in dict(q=q, dq=dq, tau=tau, singular=singular, **G).items()})
    np.savez(os.path.join(OUT, "config4_mixed_dof.npz"), **out)


if __name__ == "__main__":
    config1(); config2(); config3(); config4()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
