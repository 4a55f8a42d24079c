"""Groundwork for SURVEY.md row f-4 (batched OTG; the product refuses to enable internal OTG for now): the pinned oracle for
that row exists already -- the reference's vendored Ruckig compiles on its own (oracle/Makefile -> oracle/_ref/libotg_ref.so)
-- and tests/golden/otg_joints_reference.npz holds a trajectory it produced under the JointTask defaults (JointTask.h:38-42)
with a goal change in mid-motion.  These tests keep the fixture honest; a CUDA implementation will be held to it."""
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "otg_joints_reference.npz"))


def test_fixture_respects_the_limits_and_reaches_the_goal():
    assert np.abs(GOLD["vel"]).max() <= math.pi / 3 * (1 + 1e-12)
    assert np.abs(GOLD["acc"]).max() <= 2 * math.pi * (1 + 1e-12)
    assert np.abs(GOLD["pos"][-1] - GOLD["goals"][-1]).max() < 1e-12 and np.abs(GOLD["vel"][-1]).max() < 1e-12
    # velocity is the derivative of position, acceleration of velocity (piecewise-constant acceleration: exact trapezoid rule)
    dt = 0.001
    dp = GOLD["pos"][1:] - GOLD["pos"][:-1]
    assert np.abs(dp - 0.5 * dt * (GOLD["vel"][1:] + GOLD["vel"][:-1])).max() < 5e-6
    # phase synchronisation from rest: the first move is a straight line in joint space
    k = int(GOLD["goal_steps"][1])
    d = GOLD["goals"][0] - GOLD["q0"]
    s = (GOLD["pos"][:k] - GOLD["q0"]) @ d / (d @ d)
    assert np.abs(GOLD["pos"][:k] - GOLD["q0"] - np.outer(s, d)).max() < 1e-12


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libotg_ref.so")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_fixture_is_what_the_vendored_ruckig_produces_here():
    from tests.golden.generate_otg_reference import make_case, run_reference
    q0, goals, steps, K = make_case()
    pos, vel, acc, rc = run_reference(q0, goals, steps, K)
    assert np.array_equal(pos, GOLD["pos"]) and np.array_equal(vel, GOLD["vel"]) and np.array_equal(acc, GOLD["acc"]) and np.array_equal(rc, GOLD["rc"])
