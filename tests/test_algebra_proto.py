"""The reformulated algebra of the CUDA fast path (whitened coordinates, Householder hierarchy,
null-space joint task through the orthonormal complement, sound spectral test) against the literal
restatement of the reference, on the CPU."""
import numpy as np
import pytest

from oracle import primitives as P
from oracle.robots import make_chain
from oracle.sai_model import SaiModel, matrixRangeBasis
from tests.algebra_proto import fast_cycle, sound_nonsingular
from tests.osc_testlib import TASK_POINTS, rng_for

CASES = {
    "panda_full": ("panda", None, None),
    "panda_xyz": ("panda", [(1, 0, 0), (0, 1, 0), (0, 0, 1)], []),
    "panda_skew_dirs": ("panda", [(1, 1, 0), (0, 0, 2)], [(0, 1, 1)]),
    "rrrr_planar": ("rrrr", [(1, 0, 0), (0, 1, 0)], [(0, 0, 1)]),
    "puma_full": ("puma_like", None, None),
    "sliding_full": ("panda_sliding_base", None, None),
}


@pytest.mark.parametrize("dec", [0, 1, 2])
@pytest.mark.parametrize("case", sorted(CASES))
def test_fast_path_algebra_equals_reference_on_nonsingular_states(case, dec):
    name, dt_, dr_ = CASES[case]
    ch = make_chain(name); r = SaiModel(ch); n = ch.n
    link, pt = TASK_POINTS[name]
    checked = 0
    for trial in range(60):
        g = rng_for(trial, stream=21)
        q = ch.q_lower + (0.1 + 0.8 * g.random(n)) * (ch.q_upper - ch.q_lower); dq = g.uniform(-1, 1, n)
        r.setQ(q); r.setDq(dq); r.updateModel()
        mft = P.MotionForceTask(r, link, (np.eye(3), np.array(pt)), dt_, dr_)
        jt = P.JointTask(r)
        mft.setDynamicDecouplingType(dec); jt.setDynamicDecouplingType(dec)
        rc = P.RobotController(r, [mft, jt])
        mft.setGoalPosition(mft._current_position + g.uniform(-.05, .05, 3))
        mft.setGoalLinearVelocity(g.uniform(-.1, .1, 3)); mft.setGoalAngularAcceleration(g.uniform(-.5, .5, 3))
        jt.setGoalPosition(q + g.uniform(-.2, .2, n)); jt.setGoalAcceleration(g.uniform(-.5, .5, n))
        rc.updateControllerTaskModels(); tau = rc.computeControlTorques()
        singular = len(mft._singularity_handler._singularity_types) != 0
        Pm = mft._partial_task_projection
        bt = matrixRangeBasis(Pm[:3, :3]); br = matrixRangeBasis(Pm[3:, 3:])
        cols = []
        if np.linalg.norm(bt) > 0:
            cols += [np.concatenate([bt[:, k], np.zeros(3)]) for k in range(bt.shape[1])]
        if np.linalg.norm(br) > 0:
            cols += [np.concatenate([np.zeros(3), br[:, k]]) for k in range(br.shape[1])]
        B = np.array(cols).T
        tasks = [dict(kind="mft", B=B, J0=r.JWorldFrame(link, pt), fstar=mft._unit_mass_force, F=mft._force_related_terms, dec=dec, bie=0.1),
                 dict(kind="jt", S=np.eye(n), acc=jt._goal_acceleration, t=getattr(jt, "_pid_torques", np.zeros(n)), dec=dec, bie=0.1)]
        tau2, ok = fast_cycle(r.M(), tasks, q, dq)
        if singular:
            assert not ok, "the sound test accepted a robot the reference treats as singular"
            continue
        if ok:
            checked += 1
            assert np.abs(tau2 - tau).max() <= 1e-10 * max(np.abs(tau).max(), 1e-9)
    assert checked >= 5


def test_sound_test_never_accepts_a_singular_jacobian():
    g = rng_for(0, stream=22)
    accepted_band = 0
    for trial in range(3000):
        U, _ = np.linalg.qr(g.normal(size=(6, 6))); V, _ = np.linalg.qr(g.normal(size=(7, 6)))
        ratio = 10 ** g.uniform(-3, -0.5)
        s = np.sort(np.concatenate([[1.0, ratio], g.uniform(ratio, 1.0, 4)]))[::-1]
        J = U @ np.diag(s) @ V.T
        ok = sound_nonsingular(J, 0.06)
        if ok:
            assert ratio >= 0.06
        elif ratio > 0.0625:
            accepted_band += 1
    assert accepted_band < 30    # the rejected band above the threshold is thin
