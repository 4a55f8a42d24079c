"""ORACLE (test infrastructure only): ctypes access to oracle/_ref/libsai_ref.so -- the REFERENCE'S OWN control law.

libsai_ref.so is /root/reference/src/{RobotController,tasks/JointTask,tasks/MotionForceTask,tasks/SingularityHandler,
tasks/JointLimitAvoidanceTask,helper_modules/*}.cpp compiled where they lie, unmodified (oracle/Makefile), against stand-ins
for the two absent dependencies (oracle/eigen_standin: the part of the Eigen API those files use; oracle/saimodel_standin:
the external sai-model, i.e. kinematics/dynamics).  `kind="reference"` in the parity tests means this library: the control
law is the reference's compiled code, the model arithmetic under it is the stand-in.

`RefBatch` mirrors tests/osc_testlib.OracleBatch (N robots, one hierarchy) and hands out per-robot task proxies that forward
any method of the reference class by name (oracle/ref_wrappers/sai_ref.cpp lists them).

The library can only be BUILT where /root/reference exists; it travels to the GPU box as a built artefact.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU arms may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .robots import Chain, make_chain

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def lib_path(oriented=False):
    return os.path.join(_HERE, "_ref", "libsai_ref_orient.so" if oriented else "libsai_ref.so")


def available(oriented=False):
    if os.path.exists(lib_path(oriented)):
        return True
    if os.path.exists("/root/reference/src/RobotController.cpp"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"], stdout=subprocess.DEVNULL)
        except Exception:
            return False
    return os.path.exists(lib_path(oriented))


def load(oriented=False):
    if oriented not in _LIBS:
        if not available(oriented):
            raise RuntimeError("oracle/_ref/libsai_ref*.so is not built (it needs /root/reference)")
        lib = C.CDLL(lib_path(oriented))
        lib.sref_create.restype = C.c_void_p
        lib.sref_last_error.restype = C.c_char_p
        lib.sref_last_error.argtypes = [C.c_void_p]
        lib.sref_destroy.argtypes = [C.c_void_p]
        _LIBS[oriented] = lib
    return _LIBS[oriented]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _flat_args(args):
    out = []
    for a in args:
        if isinstance(a, (bool, np.bool_)):
            out.append(np.array([1.0 if a else 0.0]))
        else:
            out.append(np.asarray(a, dtype=np.float64).reshape(-1))
    return np.concatenate(out) if out else np.zeros(0)


class _Proxy:
    """forwards any method call to the reference object by name"""

    _SHAPES = {
        "getCurrentOrientation": (3, 3), "getGoalOrientation": (3, 3), "getDesiredOrientation": (3, 3),
        "sigmaForce": (3, 3), "sigmaPosition": (3, 3), "sigmaMoment": (3, 3), "sigmaOrientation": (3, 3),
        "posSelectionProjector": (3, 3), "oriSelectionProjector": (3, 3),
    }
    _SQUARE = ("getTaskNullspace", "getPreviousTasksNullspace", "getTaskAndPreviousNullspace", "M", "MInv", "_M_partial", "_M_partial_modified")

    def __init__(self, batch, task, robot):
        self._b, self._task, self._robot = batch, task, robot

    def _call(self, method, *args):
        a = _flat_args(args)
        out = np.zeros(512)
        r = self._b.lib.sref_call(self._b.h, C.c_int(self._task), C.c_int(self._robot), method.encode(), _p(a), C.c_int(a.size), _p(out), C.c_int(out.size))
        if r == -1:
            raise ValueError(self._b.lib.sref_last_error(self._b.h).decode())
        if r == -2:
            raise AttributeError(method)
        if r < 0:
            raise RuntimeError("sref_call(%s) failed: %d" % (method, r))
        res = out[:r].copy()
        if method in self._SHAPES:
            return res.reshape(self._SHAPES[method])
        if method in self._SQUARE:
            k = int(round(np.sqrt(r)))
            return res.reshape(k, k)
        if r == 0:
            return None
        return res

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *args: self._call(name, *args)


class _PopcView:
    def __init__(self, task):
        self._t = task

    def _s(self):
        return self._t._call("popc")

    _passivity_observer_value = property(lambda s: s._s()[0])
    _E_correction = property(lambda s: s._s()[1])
    _PO_counter = property(lambda s: int(s._s()[3]))
    _Rc = property(lambda s: s._s()[4])
    _vcl_squared_sum = property(lambda s: s._s()[5])
    window_length = property(lambda s: int(s._s()[6]))


class _SingularityView:
    def __init__(self, task):
        self._t = task

    def _s(self):
        v = self._t._call("singularity")
        k = int(v[0])
        types = [int(x) for x in v[1:1 + k]]
        c1, c2, hist, alpha, ns = v[1 + k:6 + k]
        return dict(types=types, c1=int(c1), c2=int(c2), history=int(hist), alpha=alpha, s=v[6 + k:6 + k + int(ns)].copy())

    _singularity_types = property(lambda s: s._s()["types"])
    _type_1_counter = property(lambda s: s._s()["c1"])
    _type_2_counter = property(lambda s: s._s()["c2"])
    history_length = property(lambda s: s._s()["history"])
    _alpha = property(lambda s: s._s()["alpha"])
    _svd_s = property(lambda s: s._s()["s"])
    _q_prior = property(lambda s: s._t._call("_q_prior"))
    _svd_V = property(lambda s: s._t._call("_svd_V"))


class MftProxy(_Proxy):
    """the attribute names the tests read off the numpy oracle's MotionForceTask"""
    _current_position = property(lambda s: s._call("getCurrentPosition"))
    _current_orientation = property(lambda s: s._call("getCurrentOrientation"))
    _sensed_force_control_world_frame = property(lambda s: s._call("getSensedForceControlWorldFrame"))
    _sensed_moment_control_world_frame = property(lambda s: s._call("getSensedMomentControlWorldFrame"))
    _unit_mass_force = property(lambda s: s._call("getUnitMassForce"))
    _POPC_force = property(lambda s: _PopcView(s))
    _singularity_handler = property(lambda s: _SingularityView(s))

    def setForceSensorFrame(self, link_name, T):
        R, t = T
        return self._call("setForceSensorFrame", np.asarray(R), np.asarray(t))

    def updateSensedForceAndMoment(self, f, m):
        return self._call("updateSensedForceAndMoment", np.asarray(f), np.asarray(m))

    def parametrizeForceMotionSpaces(self, dim, axis=None):
        return bool(self._call("parametrizeForceMotionSpaces", float(dim), *([] if axis is None else [np.asarray(axis)]))[0])

    def parametrizeMomentRotMotionSpaces(self, dim, axis=None):
        return bool(self._call("parametrizeMomentRotMotionSpaces", float(dim), *([] if axis is None else [np.asarray(axis)]))[0])


class JointProxy(_Proxy):
    _integrated_position_error = property(lambda s: s._call("_integrated_position_error"))


class _JlaView:
    def __init__(self, ctl):
        self._c = ctl

    _active_constraints = property(lambda s: int(s._c._call("jla_active_constraints")[0]))
    _limit_status = property(lambda s: [int(x) for x in s._c._call("jla_limit_status")])


class ControllerProxy(_Proxy):
    _joint_limit_avoidance_task = property(lambda s: _JlaView(s))


class RefBatch:
    """N robots with the same hierarchy, every one an instance of the reference's own RobotController."""

    def __init__(self, robot_name, n_robots, T_world_robot=None, oriented=False, type2_ratio_fix=True):
        self.lib = load(oriented)
        self.lib.sref_set_type2_ratio_fix(C.c_int(1 if type2_ratio_fix else 0))
        ch: Chain = make_chain(robot_name)
        self.chain, self.n, self.N = ch, ch.n, int(n_robots)
        names = list(ch.link_frames.keys())
        body = np.ascontiguousarray([ch.link_frames[k][0] for k in names], dtype=np.int32)
        lR = _c([ch.link_frames[k][1] for k in names]); lt = _c([ch.link_frames[k][2] for k in names])
        jt = np.ascontiguousarray(ch.jtype, dtype=np.int32)
        arrs = [_c(ch.axis), _c(ch.R_fix), _c(ch.t_fix), _c(ch.mass), _c(ch.com), _c(ch.inertia), _c(ch.q_lower), _c(ch.q_upper), _c(ch.dq_max), _c(ch.effort)]
        bR = bt = None
        if T_world_robot is not None:
            self._bR, self._bt = _c(T_world_robot[0]), _c(T_world_robot[1])
            bR, bt = _p(self._bR), _p(self._bt)
        self.h = C.c_void_p(self.lib.sref_create(C.c_int(ch.n), _p(jt), *[_p(a) for a in arrs], "\n".join(names).encode(), _p(body), _p(lR), _p(lt),
                                                 C.c_int(len(names)), bR, bt, C.c_int(self.N)))
        self.tasks = [[] for _ in range(self.N)]
        self._n_tasks = 0
        self.controllers = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.sref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise ValueError(self.lib.sref_last_error(self.h).decode())
        return rc

    def set_state(self, q, dq):
        q, dq = _c(q), _c(dq)
        assert q.shape == (self.N, self.n) and dq.shape == (self.N, self.n)
        self._check(self.lib.sref_set_state(self.h, _p(q), _p(dq)))

    def add_mft(self, link, compliant=None, dirs_t=None, dirs_r=None, in_compliant=False, dt=0.001, name="motion_force_task", otg=False):
        """reference default is internal OTG ON (MotionForceTask.h:67); BASELINE configs run with it off -> disabled unless otg=True"""
        cR, ct = (np.eye(3), np.zeros(3)) if compliant is None else compliant
        cR, ct = _c(cR), _c(ct)
        partial = dirs_t is not None or dirs_r is not None
        dtv = _c(dirs_t if dirs_t else np.zeros((0, 3))).reshape(-1, 3)
        drv = _c(dirs_r if dirs_r else np.zeros((0, 3))).reshape(-1, 3)
        tid = self._check(self.lib.sref_add_mft(self.h, link.encode(), _p(cR), _p(ct), C.c_int(int(partial)), _p(dtv), C.c_int(dtv.shape[0]), _p(drv),
                                                C.c_int(drv.shape[0]), C.c_int(int(in_compliant)), C.c_double(dt), name.encode()))
        out = [MftProxy(self, tid, i) for i in range(self.N)]
        if not otg:
            MftProxy(self, tid, -1).disableInternalOtg()
        for i, t in enumerate(out):
            self.tasks[i].append(t)
        return out

    def add_jt(self, S=None, dt=0.001, name="joint_task", otg=False):
        if S is None:
            tid = self._check(self.lib.sref_add_jt(self.h, None, C.c_int(self.n), C.c_double(dt), name.encode()))
        else:
            S = _c(np.atleast_2d(S))
            tid = self._check(self.lib.sref_add_jt(self.h, _p(S), C.c_int(S.shape[0]), C.c_double(dt), name.encode()))
        out = [JointProxy(self, tid, i) for i in range(self.N)]
        if not otg:
            JointProxy(self, tid, -1).disableInternalOtg()
        for i, t in enumerate(out):
            self.tasks[i].append(t)
        return out

    def finalize(self):
        self._check(self.lib.sref_finalize(self.h))
        self.controllers = [ControllerProxy(self, -1, i) for i in range(self.N)]
        self.all_controllers = ControllerProxy(self, -1, -1)

    def cycle(self, use_prev=True, n_threads=1):
        tau = np.zeros((self.N, self.n))
        self._check(self.lib.sref_cycle(self.h, _p(tau), C.c_int(int(use_prev)), C.c_int(n_threads)))
        return tau

    def step(self, q, dq, use_prev=True, n_threads=1):
        q, dq = _c(q), _c(dq)
        tau = np.zeros((self.N, self.n))
        self._check(self.lib.sref_step(self.h, _p(q), _p(dq), _p(tau), C.c_int(int(use_prev)), C.c_int(n_threads)))
        return tau

    def mft_set_goals(self, task_index, pos, ori, v, w, a, al):
        g = np.ascontiguousarray(np.concatenate([_c(pos), _c(ori).reshape(self.N, 9), _c(v), _c(w), _c(a), _c(al)], axis=1))
        self._check(self.lib.sref_mft_set_goals(self.h, C.c_int(task_index), _p(g)))

    def jt_set_goal_positions(self, task_index, pos):
        pos = _c(pos)
        self._check(self.lib.sref_jt_set_goal_positions(self.h, C.c_int(task_index), _p(pos)))

    def hardware_threads(self):
        return int(self.lib.sref_hardware_threads())
