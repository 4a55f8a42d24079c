"""ORACLE / CPU BASELINE (test infrastructure only): ctypes wrapper of oracle/cpp/osc_ref.cpp.

The C++ restatement is validated against the numpy restatement by tests/test_cpp_oracle.py and
timed by bench.py as the `cpu_baseline` / `--impl reference` arm (kind "port": the real reference
cannot be built here).  Nothing in sai_primitives_b200/ loads it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .robots import Chain, make_chain
from .sai_model import matrixRangeBasis

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboscref.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"], stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oref_create.restype = C.c_void_p
        _lib.oref_hardware_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class CppOracleBatch:
    """N robots, one hierarchy, evaluated by the C++ restatement (optionally multi-threaded)."""

    def __init__(self, robot_name: str, n_robots: int):
        self.lib = load()
        ch: Chain = make_chain(robot_name)
        self.chain = ch
        self.n = ch.n
        self.N = int(n_robots)
        jt = np.ascontiguousarray(ch.jtype, dtype=np.int32)
        arrs = [_c(ch.axis), _c(ch.R_fix), _c(ch.t_fix), _c(ch.mass), _c(ch.com), _c(ch.inertia), _c(ch.q_lower), _c(ch.q_upper), _c(ch.effort)]
        self.h = C.c_void_p(self.lib.oref_create(C.c_int(ch.n), _p(jt), *[_p(a) for a in arrs], C.c_int(self.N)))

    def close(self):
        if self.h:
            self.lib.oref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, q, dq):
        q, dq = _c(q), _c(dq)
        assert q.shape == (self.N, self.n)
        self.lib.oref_set_state(self.h, _p(q), _p(dq))

    def add_mft(self, link_name, compliant=None, dirs_t=None, dirs_r=None, in_compliant=False, dt=0.001):
        body, R_lb, t_lb = self.chain.link_frames[link_name]
        cR, ct = (np.eye(3), np.zeros(3)) if compliant is None else compliant
        P = np.eye(6)
        pr = orr = 3
        if dirs_t is not None or dirs_r is not None:
            bt = matrixRangeBasis(np.array(dirs_t, dtype=np.float64).T) if dirs_t else np.zeros((3, 1))
            br = matrixRangeBasis(np.array(dirs_r, dtype=np.float64).T) if dirs_r else np.zeros((3, 1))
            P = np.zeros((6, 6)); P[:3, :3] = bt @ bt.T; P[3:, 3:] = br @ br.T
            pr = 0 if np.linalg.norm(bt) == 0 else bt.shape[1]
            orr = 0 if np.linalg.norm(br) == 0 else br.shape[1]
        return self.lib.oref_add_mft(self.h, C.c_int(body), _p(_c(R_lb)), _p(_c(t_lb)), _p(_c(cR)), _p(_c(ct)), _p(_c(P)),
                                     C.c_int(pr), C.c_int(orr), C.c_int(1 if in_compliant else 0), C.c_double(dt))

    def add_jt(self, S=None, dt=0.001):
        if S is None:
            return self.lib.oref_add_jt(self.h, None, C.c_int(self.n), C.c_double(dt))
        S = _c(np.atleast_2d(S))
        return self.lib.oref_add_jt(self.h, _p(S), C.c_int(S.shape[0]), C.c_double(dt))

    def set_decoupling(self, task, dec, bie=0.1):
        self.lib.oref_set_decoupling(self.h, C.c_int(task), C.c_int(dec), C.c_double(bie))

    def set_options(self, gravity=False, saturation=False):
        self.lib.oref_set_options(self.h, C.c_int(1 if gravity else 0), C.c_int(1 if saturation else 0))

    def jt_set_goals(self, task, pos, vel=None, acc=None):
        pos = _c(pos)
        vel = np.zeros_like(pos) if vel is None else _c(vel)
        acc = np.zeros_like(pos) if acc is None else _c(acc)
        self.lib.oref_jt_set_goals(self.h, C.c_int(task), _p(pos), _p(vel), _p(acc))

    def jt_set_gains(self, task, kp, kv, ki, k=None):
        k = self.n if k is None else k
        a = [np.full(k, float(x)) if np.ndim(x) == 0 else _c(x) for x in (kp, kv, ki)]
        self.lib.oref_jt_set_gains(self.h, C.c_int(task), _p(a[0]), _p(a[1]), _p(a[2]))

    def mft_set_goals(self, task, pos, ori, v, w, a, al):
        g = np.concatenate([_c(pos), _c(ori).reshape(self.N, 9), _c(v), _c(w), _c(a), _c(al)], axis=1)
        g = np.ascontiguousarray(g)
        self.lib.oref_mft_set_goals(self.h, C.c_int(task), _p(g))

    def mft_get_current(self, task):
        out = np.zeros((self.N, 12))
        self.lib.oref_mft_get_current(self.h, C.c_int(task), _p(out))
        return out[:, :3].copy(), out[:, 3:].reshape(self.N, 3, 3).copy()

    def mft_force_setup(self, task, fdim=0, faxis=(0, 0, 1), mdim=0, maxis=(0, 0, 1), cl_force=False, cl_moment=False,
                        passivity=False, force_gains=None, moment_gains=None):
        fg = _p(_c(force_gains)) if force_gains is not None else None
        mg = _p(_c(moment_gains)) if moment_gains is not None else None
        self.lib.oref_mft_force_setup(self.h, C.c_int(task), C.c_int(fdim), _p(_c(faxis)), C.c_int(mdim), _p(_c(maxis)),
                                      C.c_int(int(cl_force)), C.c_int(int(cl_moment)), C.c_int(int(passivity)), fg, mg)

    def mft_set_force_goals(self, task, goal_force, goal_moment):
        g = np.ascontiguousarray(np.concatenate([_c(goal_force), _c(goal_moment)], axis=1))
        self.lib.oref_mft_set_force_goals(self.h, C.c_int(task), _p(g))

    def mft_update_sensed(self, task, force, moment):
        g = np.ascontiguousarray(np.concatenate([_c(force), _c(moment)], axis=1))
        self.lib.oref_mft_update_sensed(self.h, C.c_int(task), _p(g))

    def cycle(self, use_prev=True, n_threads=1):
        tau = np.zeros((self.N, self.n))
        self.lib.oref_cycle(self.h, _p(tau), C.c_int(1 if use_prev else 0), C.c_int(n_threads))
        return tau

    def step(self, q, dq, use_prev=True, n_threads=1):
        q, dq = _c(q), _c(dq)
        tau = np.zeros((self.N, self.n))
        self.lib.oref_step(self.h, _p(q), _p(dq), _p(tau), C.c_int(1 if use_prev else 0), C.c_int(n_threads))
        return tau

    def hardware_threads(self):
        return int(self.lib.oref_hardware_threads())
