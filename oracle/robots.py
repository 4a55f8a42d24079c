"""ORACLE (test infrastructure only) -- robot descriptions for the CPU restatement.

PARITY UNPINNED: the reference keeps its kinematics/dynamics in the un-vendored
dependency `sai-model@master` (reference CMakeLists.txt:33), which is absent
here, and the reference ships no tests.  The numbers below are the physical
parameters published in the reference's own URDF data files (cited per robot);
the merge of fixed-joint bodies into their parent follows RBDL's documented
behaviour.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
anything under oracle/.  The product (sai_primitives_b200/) has its own,
independently written model tables in C++ (csrc/host/builtin_models.cpp); a CPU
test cross-checks the two.

A robot is a serial chain.  `links` is ordered base -> tip; each entry:
  name, joint type ('fixed' | 'revolute' | 'prismatic'), xyz, rpy (joint origin
  in the parent link frame, URDF convention R = Rz(y) Ry(p) Rx(r)), axis,
  limits (lower, upper, velocity, effort), mass, com (in link frame), inertia
  (3x3 about the com, link axes).
"""
from __future__ import annotations

import math

import numpy as np

HALF_PI_URDF = 1.57079632679  # literal used by the URDFs, not pi/2


def _link(name, jtype, xyz, rpy, axis, limits, mass, com, inertia_diag):
    return dict(
        name=name, jtype=jtype, xyz=tuple(map(float, xyz)), rpy=tuple(map(float, rpy)),
        axis=tuple(map(float, axis)), limits=limits, mass=float(mass),
        com=tuple(map(float, com)), inertia=np.diag(np.asarray(inertia_diag, dtype=np.float64)),
    )


def panda_description():
    """Panda 7R + fixed end-effector.

    Source data: reference examples/15-haptic_control_impedance_type/panda_arm.urdf
    inertials :4-116, joints :118-183.
    """
    h = HALF_PI_URDF
    L = []
    # link0 is welded to the world: base of the chain (its inertia never enters M)
    L.append(_link("link0", "fixed", (0, 0, 0), (0, 0, 0), (0, 0, 1), None, 4.0, (0, 0, 0.05), (0.4, 0.4, 0.4)))
    L.append(_link("link1", "revolute", (0, 0, 0.333), (0, 0, 0), (0, 0, 1), (-2.8973, 2.8973, 2.1750, 87.0), 3.0, (0, 0, -0.07), (0.3, 0.3, 0.3)))
    L.append(_link("link2", "revolute", (0, 0, 0), (-h, 0, 0), (0, 0, 1), (-1.7628, 1.7628, 2.1750, 87.0), 3.0, (0, -0.1, 0), (0.3, 0.3, 0.3)))
    L.append(_link("link3", "revolute", (0, -0.316, 0), (h, 0, 0), (0, 0, 1), (-2.8973, 2.8973, 2.1750, 87.0), 2.0, (0.04, 0, -0.05), (0.2, 0.2, 0.2)))
    L.append(_link("link4", "revolute", (0.0825, 0, 0), (h, 0, 0), (0, 0, 1), (-3.0718, -0.0698, 2.1750, 87.0), 2.0, (-0.04, 0.05, 0), (0.2, 0.2, 0.2)))
    L.append(_link("link5", "revolute", (-0.0825, 0.384, 0), (-h, 0, 0), (0, 0, 1), (-2.8973, 2.8973, 2.6100, 12.0), 2.0, (0, 0, -0.15), (0.2, 0.2, 0.2)))
    L.append(_link("link6", "revolute", (0, 0, 0), (h, 0, 0), (0, 0, 1), (-0.0175, 3.7525, 2.6100, 12.0), 1.5, (0.06, 0, 0), (0.1, 0.1, 0.1)))
    L.append(_link("link7", "revolute", (0.088, 0, 0), (h, 0, 0), (0, 0, 1), (-2.8973, 2.8973, 2.6100, 12.0), 1.8, (0, 0, 0.17), (0.09, 0.05, 0.07)))
    L.append(_link("end-effector", "fixed", (0, 0, 0.15), (0, 0, 0), (0, 0, 1), None, 0.2, (0, 0, 0), (0.01, 0.01, 0.01)))
    return dict(name="panda", links=L)


def panda_sliding_base_description():
    """Prismatic base (axis y) + Panda.

    Source data: reference examples/06-partial_joint_task/panda_arm_sliding_base.urdf
    :171-233.  The joint0 origin literal in that file is malformed ("0 0 0.-75");
    it is taken as (0, 0, 0), which is what atof() yields.
    """
    d = panda_description()
    links = d["links"]
    slider = _link("slider_link", "fixed", (0, 0, 0), (0, 0, 0), (0, 0, 1), None, 4.0, (0, 0, 0.05), (0.4, 0.4, 0.4))
    link0 = dict(links[0])
    link0.update(jtype="prismatic", axis=(0.0, 1.0, 0.0), limits=(-1.0, 1.0, 2.0, 150.0))
    return dict(name="panda_sliding_base", links=[slider, link0] + links[1:])


def rrrr_description():
    """Planar 4R arm.

    Source data: reference examples/11-planar_robot_controller/rrrrbot.urdf
    inertials :5-127, joints :135-166.
    """
    I = (0.084167, 0.083467, 0.000967)
    lim = (-2.9, 2.9, 1.7104, 176.0)
    L = [_link("link0", "fixed", (0, 0, 0), (0, 0, 0), (0, 0, 1), None, 1.0, (0, 0, 0), I)]
    L.append(_link("link1", "revolute", (0, 0, 0), (0, 0, 0), (0, 0, 1), lim, 1.0, (0.25, 0, 0), I))
    for k in (2, 3, 4):
        L.append(_link("link%d" % k, "revolute", (0.5, 0, 0), (0, 0, 0), (0, 0, 1), lim, 1.0, (0.25, 0, 0), I))
    return dict(name="rrrr", links=L)


def puma_like_description():
    """A PUMA-560-like 6R arm AUTHORED BY THIS REPO (the reference's puma.urdf
    lives in sai-model and is not available).  Geometry is
    the classic shoulder/elbow/spherical-wrist layout; parity for it is against
    this oracle only."""
    h = math.pi / 2
    L = [_link("base", "fixed", (0, 0, 0), (0, 0, 0), (0, 0, 1), None, 10.0, (0, 0, 0.3), (1.0, 1.0, 0.5))]
    L.append(_link("shoulder", "revolute", (0, 0, 0.66), (0, 0, 0), (0, 0, 1), (-2.79, 2.79, 2.0, 100.0), 8.0, (0, 0, -0.1), (0.30, 0.30, 0.35)))
    L.append(_link("upper_arm", "revolute", (0, 0.15, 0), (-h, 0, 0), (0, 0, 1), (-3.92, 0.78, 2.0, 100.0), 12.0, (0.20, 0, 0.05), (0.13, 0.52, 0.54)))
    L.append(_link("forearm", "revolute", (0.4318, 0, 0), (0, 0, 0), (0, 0, 1), (-0.78, 3.92, 2.0, 80.0), 4.8, (0.02, -0.15, 0), (0.066, 0.0125, 0.086)))
    L.append(_link("wrist1", "revolute", (0.0203, -0.4331, 0), (h, 0, 0), (0, 0, 1), (-1.92, 2.97, 3.0, 20.0), 0.82, (0, 0, -0.02), (0.0018, 0.0018, 0.0013)))
    L.append(_link("wrist2", "revolute", (0, 0, 0), (-h, 0, 0), (0, 0, 1), (-1.74, 1.74, 3.0, 20.0), 0.34, (0, 0, 0), (0.0003, 0.0003, 0.0004)))
    L.append(_link("wrist3", "revolute", (0, 0, 0), (h, 0, 0), (0, 0, 1), (-4.64, 4.64, 3.0, 20.0), 0.09, (0, 0, 0.03), (0.00015, 0.00015, 0.00004)))
    L.append(_link("end-effector", "fixed", (0, 0, 0.056), (0, 0, 0), (0, 0, 1), None, 0.05, (0, 0, 0), (0.00001, 0.00001, 0.00001)))
    return dict(name="puma_like", links=L)


DESCRIPTIONS = {
    "panda": panda_description,
    "panda_sliding_base": panda_sliding_base_description,
    "rrrr": rrrr_description,
    "puma_like": puma_like_description,
}


def rpy_to_matrix(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]], dtype=np.float64)
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=np.float64)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], dtype=np.float64)
    return Rz @ Ry @ Rx


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


class Chain:
    """Movable-joint chain after merging fixed-joint bodies into their parent
    (RBDL URDF-reader behaviour).  Body i is moved by
    joint i; body -1 is the world."""

    def __init__(self, desc):
        self.name = desc["name"]
        n = sum(1 for l in desc["links"] if l["jtype"] != "fixed")
        self.n = n
        self.jtype = np.zeros(n, dtype=np.int32)       # 0 revolute, 1 prismatic
        self.axis = np.zeros((n, 3))
        self.R_fix = np.zeros((n, 3, 3))               # parent body frame -> joint frame (q = 0)
        self.t_fix = np.zeros((n, 3))
        self.mass = np.zeros(n)
        self.com = np.zeros((n, 3))                    # in body frame
        self.inertia = np.zeros((n, 3, 3))             # about com, body axes
        self.q_lower = np.zeros(n); self.q_upper = np.zeros(n)
        self.dq_max = np.zeros(n); self.effort = np.zeros(n)
        self.link_frames = {}                          # link name -> (body, R_in_body, t_in_body)
        self.joint_names = []

        body = -1
        # pose of the current URDF link frame expressed in the current body frame
        R_lb, t_lb = np.eye(3), np.zeros(3)
        # accumulate (mass, first moment, second moment about body origin)
        acc = None
        for l in desc["links"]:
            R_j = rpy_to_matrix(l["rpy"]); t_j = np.asarray(l["xyz"])
            if l["jtype"] == "fixed":
                R_lb, t_lb = R_lb @ R_j, t_lb + R_lb @ t_j
            else:
                if body >= 0:
                    self._finish_body(body, acc)
                body += 1
                self.jtype[body] = 0 if l["jtype"] == "revolute" else 1
                a = np.asarray(l["axis"]); self.axis[body] = a / np.linalg.norm(a)
                self.R_fix[body] = R_lb @ R_j
                self.t_fix[body] = t_lb + R_lb @ t_j
                lo, hi, vel, eff = l["limits"]
                self.q_lower[body], self.q_upper[body] = lo, hi
                self.dq_max[body], self.effort[body] = vel, eff
                self.joint_names.append("joint_" + l["name"])
                R_lb, t_lb = np.eye(3), np.zeros(3)
                acc = [0.0, np.zeros(3), np.zeros((3, 3))]
            self.link_frames[l["name"]] = (body, R_lb.copy(), t_lb.copy())
            if body >= 0:
                m = l["mass"]; c = t_lb + R_lb @ np.asarray(l["com"])
                Ic = R_lb @ l["inertia"] @ R_lb.T
                acc[0] += m
                acc[1] += m * c
                acc[2] += Ic + m * (np.dot(c, c) * np.eye(3) - np.outer(c, c))
        if body >= 0:
            self._finish_body(body, acc)

    def _finish_body(self, b, acc):
        m, h, Io = acc
        c = h / m
        self.mass[b] = m
        self.com[b] = c
        self.inertia[b] = Io - m * (np.dot(c, c) * np.eye(3) - np.outer(c, c))


def make_chain(name: str) -> Chain:
    return Chain(DESCRIPTIONS[name]())
