"""TEST INFRASTRUCTURE (oracle): batched-simulation side of the control loop, SURVEY.md row f-1.

The reference examples close the loop with sai-simulation (`sim->setJointTorques(tau); sim->integrate()`,
examples/05-using_robot_controller/05-using_robot_controller.cpp:225-231), which is an external dependency that is not in
/root/reference: PARITY UNPINNED.  What is restated here is the textbook rigid-body forward dynamics of the same serial chain
the controller uses,  ddq = M(q)^-1 (tau - b(q, dq) - g(q)),  integrated with the semi-implicit Euler scheme SURVEY.md f-1
names (dq += ddq dt, then q += dq dt).  b + g comes from the recursive Newton-Euler algorithm in the robot base frame;
tests/test_oracle_model.py checks it against two independent statements: the gravity vector from the Jacobians
(SaiModel.jointGravityVector) and the Coriolis/centrifugal forces from the Christoffel symbols of a finite-difference dM/dq.
"""
from __future__ import annotations

import numpy as np

from .sai_model import SaiModel


def bias_forces(model: SaiModel) -> np.ndarray:
    """b(q, dq) + g(q): joint torques that hold ddq = 0 (recursive Newton-Euler, base frame)."""
    c = model.chain
    n = c.n
    dq = model.dq()
    g_base = model._R_wr.T @ model._gravity_world
    w = np.zeros(3); al = np.zeros(3); ap = -g_base          # base: at rest, accelerating "upwards" with -g
    p_prev = np.zeros(3)
    W = np.zeros((n, 3)); AL = np.zeros((n, 3)); AP = np.zeros((n, 3))
    for k in range(n):
        a, p = model._ax[k], model._pb[k]
        r = p - p_prev
        ap = ap + np.cross(al, r) + np.cross(w, np.cross(w, r))
        if c.jtype[k] == 0:
            al = al + np.cross(w, a) * dq[k]
            w = w + a * dq[k]
        else:
            ap = ap + 2.0 * np.cross(w, a) * dq[k]
        W[k], AL[k], AP[k] = w, al, ap
        p_prev = p
    f = np.zeros(3); nm = np.zeros(3)                         # force / moment (about the joint origin) transmitted through joint k
    tau = np.zeros(n)
    p_next = None
    for k in range(n - 1, -1, -1):
        R = model._Rb[k]
        cw = R @ c.com[k]
        Iw = R @ c.inertia[k] @ R.T
        ac = AP[k] + np.cross(AL[k], cw) + np.cross(W[k], np.cross(W[k], cw))
        F = c.mass[k] * ac
        Nc = Iw @ AL[k] + np.cross(W[k], Iw @ W[k])
        if p_next is not None:
            nm = nm + np.cross(p_next - model._pb[k], f)
        nm = nm + Nc + np.cross(cw, F)
        f = f + F
        tau[k] = model._ax[k] @ (nm if c.jtype[k] == 0 else f)
        p_next = model._pb[k]
    return tau


def forward_dynamics(model: SaiModel, tau: np.ndarray) -> np.ndarray:
    return np.linalg.solve(model.M(), np.asarray(tau, dtype=np.float64) - bias_forces(model))


def integrate(model: SaiModel, tau: np.ndarray, dt: float, substeps: int = 1) -> None:
    """semi-implicit Euler with the torque held over the step; updates the model's q, dq in place"""
    for _ in range(substeps):
        ddq = forward_dynamics(model, tau)
        dq = model.dq() + ddq * dt
        q = model.q() + dq * dt
        model.setQ(q); model.setDq(dq); model.updateModel()
