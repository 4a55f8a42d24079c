"""the product's OTG core (sai_primitives_b200/csrc/osc_otg.h -- the header the CUDA kernels include, built
for the host by this test) against THE REFERENCE'S OWN CODE on the CPU:
  * tests/golden/otg_joints_reference.npz: a trajectory of the reference's vendored Ruckig under the JointTask defaults with a
    goal change in mid-motion (generator: tests/golden/generate_otg_reference.py);
  * where oracle/_ref/libsai_ref.so exists: JointTask with its internal OTG left ON (the reference's default, JointTask.h:38),
    i.e. the reference's OTG_joints.cpp + Ruckig compiled in place, on random goal sequences incl. non-collinear re-targeting,
    per-joint limits and re-initialisation."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TOL = 1e-10      # both sides evaluate the same closed forms in FP64; positions are O(1) rad
P = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    out = tmp_path_factory.mktemp("otg") / "libotg_probe.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", str(out), os.path.join(HERE, "cpp", "otg_host_probe.cpp")])
    lib = C.CDLL(str(out))
    lib.otgp_create.restype = C.c_void_p
    lib.otgp_create.argtypes = [C.c_int, C.c_double, P]
    lib.otgp_destroy.argtypes = [C.c_void_p]
    lib.otgp_set_limits.argtypes = [C.c_void_p, P, P]
    lib.otgp_set_goal.argtypes = [C.c_void_p, P, P]
    lib.otgp_reinitialize.argtypes = [C.c_void_p, P]
    lib.otgp_update.argtypes = [C.c_void_p, P, P, P]
    lib.otgp_update.restype = C.c_int
    return lib


def _p(a):
    return a.ctypes.data_as(P)


class ProbeOtg:
    def __init__(self, lib, q0, dt, vmax, amax):
        self.lib, self.k = lib, q0.size
        q0 = np.ascontiguousarray(q0, dtype=np.float64)
        self.h = lib.otgp_create(self.k, dt, _p(q0))
        v = np.ascontiguousarray(np.broadcast_to(vmax, (self.k,)), dtype=np.float64); a = np.ascontiguousarray(np.broadcast_to(amax, (self.k,)), dtype=np.float64)
        lib.otgp_set_limits(self.h, _p(v), _p(a))

    def set_goal(self, pos, vel=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64); vel = np.zeros(self.k) if vel is None else np.ascontiguousarray(vel, dtype=np.float64)
        self.lib.otgp_set_goal(self.h, _p(pos), _p(vel))

    def reinitialize(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        self.lib.otgp_reinitialize(self.h, _p(pos))

    def update(self):
        p, v, a = np.zeros(self.k), np.zeros(self.k), np.zeros(self.k)
        flags = self.lib.otgp_update(self.h, _p(p), _p(v), _p(a))
        return p, v, a, flags


def test_core_reproduces_the_reference_ruckig_trajectory(probe):
    d = np.load(os.path.join(HERE, "golden", "otg_joints_reference.npz"))
    o = ProbeOtg(probe, d["q0"], 0.001, math.pi / 3, 2 * math.pi)
    K = d["pos"].shape[0]
    worst = 0.0
    for k in range(K):
        for g, s in zip(d["goals"], d["goal_steps"]):
            if s == k:
                o.set_goal(g)
        p, v, a, flags = o.update()
        worst = max(worst, np.abs(p - d["pos"][k]).max(), np.abs(v - d["vel"][k]).max(), np.abs(a - d["acc"][k]).max() * 1e-3)
        assert worst < TOL, (k, worst)
        assert bool(flags & 1) == bool(d["rc"][k] == 1 and np.linalg.norm(d["vel"][k]) < 1e-3) or k < K - 1
    assert flags & 1      # goal reached at the end


def _reference_available():
    from oracle import sai_ref
    return sai_ref.available(oriented=False)


@pytest.mark.skipif(not _reference_available(), reason="oracle/_ref/libsai_ref.so needs /root/reference to be built")
@pytest.mark.parametrize("robot_name,per_joint_limits", [("panda", False), ("panda", True), ("rrrr", False), ("panda_sliding_base", True)])
def test_core_matches_the_reference_joint_task_otg(probe, robot_name, per_joint_limits):
    """desired position / velocity / acceleration of the reference's JointTask with internal OTG on"""
    from oracle.sai_ref import RefBatch
    from tests.osc_testlib import sample_states
    N = 6
    q, dq, _ = sample_states(robot_name, N)
    n = q.shape[1]
    rb = RefBatch(robot_name, N); rb.set_state(q, dq)
    jt = rb.add_jt(otg=True); rb.finalize()
    rng = np.random.default_rng(11)
    if per_joint_limits:
        vmax = rng.uniform(0.4, 1.5, n); amax = rng.uniform(2.0, 8.0, n)
        for t in jt:
            t.enableInternalOtgAccelerationLimited(vmax, amax)
    else:
        vmax, amax = math.pi / 3, 2 * math.pi      # JointTask.h:40-41
    probes = [ProbeOtg(probe, q[i], 0.001, vmax, amax) for i in range(N)]
    K = 1800
    events = {0: "goal", 300: "goal", 420: "goal_scaled", 1100: "reinit", 1200: "goal"}
    worst = 0.0
    goal = q.copy()
    for k in range(K):
        ev = events.get(k)
        if ev == "goal":
            goal = q + rng.uniform(-0.8, 0.8, (N, n))
        elif ev == "goal_scaled":      # along the current direction of motion: stays collinear -> phase synchronisation again
            goal = q + 1.3 * (goal - q)
        if ev in ("goal", "goal_scaled"):
            for i in range(N):
                jt[i].setGoalPosition(goal[i]); probes[i].set_goal(goal[i])
        if ev == "reinit":
            for i in range(N):
                jt[i].reInitializeTask(); probes[i].reinitialize(q[i])
            goal = q.copy()
        rb.cycle()
        for i in range(N):
            p, v, a, flags = probes[i].update()
            worst = max(worst, np.abs(p - jt[i].getDesiredPosition()).max(), np.abs(v - jt[i].getDesiredVelocity()).max(),
                        np.abs(a - jt[i].getDesiredAcceleration()).max() * 1e-3)
        assert worst < TOL, (k, worst)
    assert worst < TOL


@pytest.mark.skipif(not _reference_available(), reason="oracle/_ref/libsai_ref.so needs /root/reference to be built")
def test_core_matches_the_reference_motion_force_task_otg(probe):
    """desired pose / velocities / accelerations of the reference's MotionForceTask with internal OTG on (the default,
    MotionForceTask.h:67-74): OTG_6dof_cartesian.cpp + Ruckig<6> compiled in place; goal changes in mid-motion, a small change below
    the reference's 1e-3 relative threshold (ignored by both), re-initialisation"""
    from oracle.sai_ref import RefBatch
    from tests.osc_testlib import TASK_POINTS, rot_exp, sample_states
    probe.otgc_create.restype = C.c_void_p
    probe.otgc_create.argtypes = [C.c_double, P, P, C.c_double, C.c_double, C.c_double, C.c_double]
    probe.otgc_set_goal.argtypes = [C.c_void_p, P, P, P, P]
    probe.otgc_reinitialize.argtypes = [C.c_void_p, P, P]
    probe.otgc_update.argtypes = [C.c_void_p, P]
    probe.otgc_update.restype = C.c_int
    N = 5
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.1)
    link, pt = TASK_POINTS["panda"]
    rb = RefBatch("panda", N); rb.set_state(q, dq)
    mft = rb.add_mft(link, (np.eye(3), np.array(pt)), otg=True); rb.finalize()
    x0 = np.array([t.getCurrentPosition() for t in mft]); R0 = np.array([t.getCurrentOrientation() for t in mft])
    hs = [probe.otgc_create(0.001, _p(np.ascontiguousarray(x0[i])), _p(np.ascontiguousarray(R0[i])), 0.3, 2.0, math.pi / 3, 2 * math.pi) for i in range(N)]
    rng = np.random.default_rng(5)
    K = 1500
    events = {0: "goal", 250: "goal", 400: "tiny", 700: "goal_vel", 1000: "reinit", 1100: "goal"}
    gx, gR = x0.copy(), R0.copy()
    gv, gw = np.zeros((N, 3)), np.zeros((N, 3))
    worst = 0.0
    out = np.zeros(24)
    for k in range(K):
        ev = events.get(k)
        if ev in ("goal", "goal_vel"):
            gx = x0 + rng.uniform(-0.15, 0.15, (N, 3)); gR = np.array([R0[i] @ rot_exp(rng.uniform(-0.6, 0.6, 3)) for i in range(N)])
            gv = rng.uniform(-0.05, 0.05, (N, 3)) if ev == "goal_vel" else np.zeros((N, 3))
            gw = rng.uniform(-0.1, 0.1, (N, 3)) if ev == "goal_vel" else np.zeros((N, 3))
        elif ev == "tiny":
            gx = gx * (1.0 + 2e-4)
        elif ev == "reinit":
            for i in range(N):
                mft[i].reInitializeTask(); probe.otgc_reinitialize(hs[i], _p(np.ascontiguousarray(x0[i])), _p(np.ascontiguousarray(R0[i])))
            gx, gR, gv, gw = x0.copy(), R0.copy(), np.zeros((N, 3)), np.zeros((N, 3))
        for i in range(N):
            t = mft[i]
            t.setGoalPosition(gx[i]); t.setGoalOrientation(gR[i]); t.setGoalLinearVelocity(gv[i]); t.setGoalAngularVelocity(gw[i])
            probe.otgc_set_goal(hs[i], _p(np.ascontiguousarray(gx[i])), _p(np.ascontiguousarray(gv[i])), _p(np.ascontiguousarray(gR[i])), _p(np.ascontiguousarray(gw[i])))
        rb.cycle()
        for i in range(N):
            probe.otgc_update(hs[i], _p(out))
            t = mft[i]
            ref = np.concatenate([t.getDesiredPosition(), t.getDesiredOrientation().reshape(-1), t.getDesiredLinearVelocity(), t.getDesiredAngularVelocity(),
                                  1e-3 * t.getDesiredLinearAcceleration(), 1e-3 * t.getDesiredAngularAcceleration()])
            mine = out.copy(); mine[18:] *= 1e-3
            worst = max(worst, np.abs(mine - ref).max())
        assert worst < 1e-9, (k, worst)
