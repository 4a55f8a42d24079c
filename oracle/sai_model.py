"""ORACLE (test infrastructure only) -- restatement of the sai-model calls the
reference's hot path makes.

PARITY UNPINNED.  sai-model (github manips-sai-org/sai-model, taken @master by
the reference's CI, .github/actions/build-repo/action.yml:20-23) is NOT in
/root/reference and not installable here.  This file restates the published,
standard algorithms behind each call the reference makes (call sites listed in
SURVEY.md section 8c) with numpy/LAPACK float64:

  updateModel             forward kinematics + mass matrix + inverse
  J / JWorldFrame         6 x n point Jacobian, linear rows first
                          (row order evidenced by MotionForceTask.cpp:293-298)
  position/rotation[...]  point position / frame rotation
  jointGravityVector      sum_k  J_v,com_k^T (-m_k g)
  operationalSpaceMatrices  Lambda=(J M^-1 J^T)^-1, Jbar=M^-1 J^T Lambda, N=I-Jbar J
  matrixRangeBasis        SVD range basis with relative tolerance 1e-3
  orientationError        -1/2 sum_i Rc[:,i] x Rd[:,i]
  computePseudoInverse    SVD Moore-Penrose

The mass matrix is computed from its DEFINITION (sum over bodies of
m Jv^T Jv + Jw^T I Jw), deliberately not by the CRBA recursion the CUDA kernels
use, so the two are independent derivations of the same quantity.
"""
from __future__ import annotations

import numpy as np

from .robots import Chain, skew


def rot_axis_angle(axis, angle):
    """Rodrigues rotation about a unit axis."""
    K = skew(axis)
    return np.eye(3) + np.sin(angle) * K + (1.0 - np.cos(angle)) * (K @ K)


class OpSpaceMatrices:
    def __init__(self, J, Lambda, Jbar, N):
        self.J, self.Lambda, self.Jbar, self.N = J, Lambda, Jbar, N


class JointLimit:
    def __init__(self, name, index, lo, hi, vel, eff):
        self.joint_name, self.joint_index = name, index
        self.position_lower, self.position_upper = lo, hi
        self.velocity, self.effort = vel, eff


class SaiModel:
    def __init__(self, chain: Chain, T_world_robot=None, gravity=(0.0, 0.0, -9.81)):
        self.chain = chain
        n = chain.n
        self._q = np.zeros(n)
        self._dq = np.zeros(n)
        self._R_wr = np.eye(3) if T_world_robot is None else np.asarray(T_world_robot[0], dtype=np.float64)
        self._t_wr = np.zeros(3) if T_world_robot is None else np.asarray(T_world_robot[1], dtype=np.float64)
        self._gravity_world = np.asarray(gravity, dtype=np.float64)
        self._M = np.eye(n)
        self._M_inv = np.eye(n)
        self.updateModel()

    # ---- state ----
    def dof(self):
        return self.chain.n

    def q(self):
        return self._q.copy()

    def dq(self):
        return self._dq.copy()

    def setQ(self, q):
        q = np.asarray(q, dtype=np.float64)
        assert q.shape == (self.chain.n,)
        self._q = q.copy()

    def setDq(self, dq):
        dq = np.asarray(dq, dtype=np.float64)
        assert dq.shape == (self.chain.n,)
        self._dq = dq.copy()

    # ---- kinematics ----
    def updateKinematics(self):
        c = self.chain
        n = c.n
        self._Rb = np.zeros((n, 3, 3))   # body frame orientation in robot base frame
        self._pb = np.zeros((n, 3))      # body frame origin in robot base frame
        self._ax = np.zeros((n, 3))      # joint axis in robot base frame
        R, p = np.eye(3), np.zeros(3)
        for i in range(n):
            p = p + R @ c.t_fix[i]
            R = R @ c.R_fix[i]
            if c.jtype[i] == 0:
                R = R @ rot_axis_angle(c.axis[i], self._q[i])
            else:
                p = p + R @ c.axis[i] * self._q[i]
            self._Rb[i], self._pb[i] = R, p
            self._ax[i] = R @ c.axis[i]

    def updateModel(self):
        self.updateKinematics()
        c = self.chain
        n = c.n
        M = np.zeros((n, n))
        for k in range(n):
            pc = self._pb[k] + self._Rb[k] @ c.com[k]
            Jk = self._jacobian_body_point(k, pc)
            Iw = self._Rb[k] @ c.inertia[k] @ self._Rb[k].T
            M += c.mass[k] * Jk[:3].T @ Jk[:3] + Jk[3:].T @ Iw @ Jk[3:]
        self._M = 0.5 * (M + M.T)
        self._M_inv = np.linalg.inv(self._M)   # Eigen dynamic .inverse() is PartialPivLU too

    def _jacobian_body_point(self, body, p_base):
        n = self.chain.n
        J = np.zeros((6, n))
        for i in range(body + 1):
            a = self._ax[i]
            if self.chain.jtype[i] == 0:
                J[:3, i] = np.cross(a, p_base - self._pb[i])
                J[3:, i] = a
            else:
                J[:3, i] = a
        return J

    def _link(self, link_name):
        if link_name not in self.chain.link_frames:
            raise ValueError("link [%s] does not exist" % link_name)
        return self.chain.link_frames[link_name]

    def position(self, link_name, pos_in_link=(0.0, 0.0, 0.0)):
        b, R_lb, t_lb = self._link(link_name)
        local = t_lb + R_lb @ np.asarray(pos_in_link, dtype=np.float64)
        if b < 0:
            return local
        return self._pb[b] + self._Rb[b] @ local

    def rotation(self, link_name, rot_in_link=None):
        b, R_lb, _ = self._link(link_name)
        Rl = R_lb if rot_in_link is None else R_lb @ np.asarray(rot_in_link, dtype=np.float64)
        return Rl if b < 0 else self._Rb[b] @ Rl

    def positionInWorld(self, link_name, pos_in_link=(0.0, 0.0, 0.0)):
        return self._t_wr + self._R_wr @ self.position(link_name, pos_in_link)

    def rotationInWorld(self, link_name, rot_in_link=None):
        return self._R_wr @ self.rotation(link_name, rot_in_link)

    def transformInWorld(self, link_name):
        return self.rotationInWorld(link_name), self.positionInWorld(link_name)

    def J(self, link_name, pos_in_link=(0.0, 0.0, 0.0)):
        b, _, _ = self._link(link_name)
        if b < 0:
            return np.zeros((6, self.chain.n))
        return self._jacobian_body_point(b, self.position(link_name, pos_in_link))

    def JWorldFrame(self, link_name, pos_in_link=(0.0, 0.0, 0.0)):
        J = self.J(link_name, pos_in_link)
        return np.vstack([self._R_wr @ J[:3], self._R_wr @ J[3:]])

    # ---- dynamics ----
    def M(self):
        return self._M.copy()

    def MInv(self):
        return self._M_inv.copy()

    def jointGravityVector(self):
        c = self.chain
        g_base = self._R_wr.T @ self._gravity_world
        tau = np.zeros(c.n)
        for k in range(c.n):
            pc = self._pb[k] + self._Rb[k] @ c.com[k]
            Jk = self._jacobian_body_point(k, pc)
            tau += Jk[:3].T @ (-c.mass[k] * g_base)
        return tau

    def operationalSpaceMatrices(self, task_jacobian):
        J = np.asarray(task_jacobian, dtype=np.float64)
        n = self.chain.n
        inv_inertia = J @ self._M_inv @ J.T
        # Eigen: inv_inertia.llt().solve(Identity)
        Lc = np.linalg.cholesky(0.5 * (inv_inertia + inv_inertia.T))
        Linv = np.linalg.solve(Lc, np.eye(J.shape[0]))
        Lambda = Linv.T @ Linv
        Jbar = self._M_inv @ J.T @ Lambda
        N = np.eye(n) - Jbar @ J
        return OpSpaceMatrices(J, Lambda, Jbar, N)

    def nullspaceMatrix(self, task_jacobian):
        return self.operationalSpaceMatrices(task_jacobian).N

    def jointLimits(self):
        c = self.chain
        return [JointLimit(c.joint_names[i], i, c.q_lower[i], c.q_upper[i], c.dq_max[i], c.effort[i]) for i in range(c.n)]

    def kineticEnergy(self):
        """Independent check value: sum over bodies of 1/2 m v_c^2 + 1/2 w^T I w."""
        c = self.chain
        e = 0.0
        for k in range(c.n):
            pc = self._pb[k] + self._Rb[k] @ c.com[k]
            Jk = self._jacobian_body_point(k, pc)
            v = Jk[:3] @ self._dq
            w = Jk[3:] @ self._dq
            Iw = self._Rb[k] @ c.inertia[k] @ self._Rb[k].T
            e += 0.5 * c.mass[k] * v @ v + 0.5 * w @ Iw @ w
        return e


# ---------------- free functions of sai-model used by the reference ----------------

def matrixRangeBasis(matrix, tolerance=1e-3):
    """SURVEY.md Appendix B; zero-matrix convention evidenced by
    JointTask.cpp:234 and MotionForceTask.cpp:151-152."""
    A = np.atleast_2d(np.asarray(matrix, dtype=np.float64))
    rows = A.shape[0]
    if np.linalg.norm(A) < tolerance:
        return np.zeros((rows, 1))
    U, s, _ = np.linalg.svd(A, full_matrices=False)
    if s[0] < tolerance:
        return np.zeros((rows, 1))
    task_dof = min(A.shape)
    for i in range(len(s) - 1, 0, -1):
        if s[i] / s[0] < tolerance:
            task_dof -= 1
        else:
            break
    if task_dof == rows:
        return np.eye(rows)
    return U[:, :task_dof].copy()


def orientationError(desired_orientation, current_orientation):
    Rd = np.asarray(desired_orientation, dtype=np.float64)
    Rc = np.asarray(current_orientation, dtype=np.float64)
    for R, nm in ((Rd, "desired"), (Rc, "current")):
        if np.linalg.norm(R.T @ R - np.eye(3)) > 1e-4 or abs(np.linalg.det(R) - 1.0) > 1e-4:
            raise ValueError("%s orientation is not a valid rotation matrix" % nm)
    e = np.zeros(3)
    for i in range(3):
        e += np.cross(Rc[:, i], Rd[:, i])
    return -0.5 * e


def computePseudoInverse(matrix, tolerance=1e-6):
    A = np.asarray(matrix, dtype=np.float64)
    U, s, Vt = np.linalg.svd(A, full_matrices=False)
    s_inv = np.array([1.0 / x if x > tolerance else 0.0 for x in s])
    return Vt.T @ np.diag(s_inv) @ U.T
