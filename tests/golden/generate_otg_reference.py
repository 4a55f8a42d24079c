"""Groundwork for SURVEY.md row f-4 (batched OTG, not built in the product yet): golden trajectories from the REFERENCE'S
vendored Ruckig (oracle/_ref/libotg_ref.so, oracle/Makefile) under the JointTask defaults (JointTask.h:38-42: OTG on,
acceleration-limited, max velocity pi/3, max acceleration 2 pi, phase synchronisation OTG_joints.cpp:24).
    make -C oracle && python tests/golden/generate_otg_reference.py
"""
import ctypes as C
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def load_ref():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libotg_ref.so"))
    P = C.POINTER(C.c_double)
    lib.otg_ref_create.restype = C.c_void_p
    lib.otg_ref_create.argtypes = [C.c_int, C.c_double, P]
    lib.otg_ref_destroy.argtypes = [C.c_void_p]
    lib.otg_ref_set_limits.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    lib.otg_ref_set_goal.argtypes = [C.c_void_p, P, P]
    lib.otg_ref_update.argtypes = [C.c_void_p, P, P, P]
    lib.otg_ref_update.restype = C.c_int
    return lib


def run_reference(q0, goals, goal_steps, K, dt=0.001, vmax=math.pi / 3, amax=2 * math.pi, jmax=0.0):
    lib = load_ref()
    P = C.POINTER(C.c_double)
    n = q0.size
    q0 = np.ascontiguousarray(q0, dtype=np.float64)
    h = lib.otg_ref_create(n, dt, q0.ctypes.data_as(P))
    lib.otg_ref_set_limits(h, vmax, amax, jmax)
    pos = np.zeros((K, n)); vel = np.zeros((K, n)); acc = np.zeros((K, n)); rc = np.zeros(K, dtype=np.int32)
    zero = np.zeros(n)
    for k in range(K):
        for g, s in zip(goals, goal_steps):
            if s == k:
                g = np.ascontiguousarray(g)
                lib.otg_ref_set_goal(h, g.ctypes.data_as(P), zero.ctypes.data_as(P))
        p = np.zeros(n); v = np.zeros(n); a = np.zeros(n)
        rc[k] = lib.otg_ref_update(h, p.ctypes.data_as(P), v.ctypes.data_as(P), a.ctypes.data_as(P))
        pos[k], vel[k], acc[k] = p, v, a
    lib.otg_ref_destroy(h)
    return pos, vel, acc, rc


def make_case(seed=7):
    g = np.random.default_rng(seed)
    q0 = g.uniform(-1, 1, 7)
    goals = [q0 + g.uniform(-0.6, 0.6, 7), q0 + g.uniform(-0.3, 0.3, 7), q0 + g.uniform(-1.0, 1.0, 7)]
    steps = [0, 350, 900]          # the second goal arrives while the first move is under way
    return q0, goals, steps, 2600


if __name__ == "__main__":
    q0, goals, steps, K = make_case()
    pos, vel, acc, rc = run_reference(q0, goals, steps, K)
    print("max |v| %.4f (limit %.4f), max |a| %.4f (limit %.4f), final error %.2e, finished at step %d"
          % (np.abs(vel).max(), math.pi / 3, np.abs(acc).max(), 2 * math.pi, np.abs(pos[-1] - goals[-1]).max(), int(np.argmax(rc[steps[-1]:] == 1)) + steps[-1]))
    np.savez_compressed(os.path.join(HERE, "otg_joints_reference.npz"), q0=q0, goals=np.array(goals), goal_steps=np.array(steps), pos=pos, vel=vel, acc=acc, rc=rc)
