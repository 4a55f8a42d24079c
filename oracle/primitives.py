"""ORACLE (test infrastructure only) -- numpy float64 restatement of the
reference control law, statement by statement.

PARITY UNPINNED: the reference has no tests / golden vectors for this path and
cannot be built here (needs Eigen + sai-model); this file follows the reference
SOURCE line by line instead, each function citing the file:line it follows.
Exception: POPCExplicitForceControl below is pinned -- the reference's own
POPCExplicitForceControl.cpp compiles against a small stand-in for the Eigen types
it uses (oracle/Makefile, oracle/_ref) and tests/test_popc_reference.py holds this
class to its outputs (tests/golden/popc_reference.npz) to 1e-12.
Internal OTG (Ruckig) is excluded (BASELINE.json north_star) and defaults OFF.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module; nothing in sai_primitives_b200/ does.
"""
from __future__ import annotations

import math
from collections import deque

import numpy as np

from .sai_model import SaiModel, computePseudoInverse, matrixRangeBasis, orientationError

# reference src/helper_modules/SaiPrimitivesCommonDefinitions.h:14-20
FULL_DYNAMIC_DECOUPLING = 0
BOUNDED_INERTIA_ESTIMATES = 1
IMPEDANCE = 2

# reference src/tasks/SingularityHandler.h:25-29
NO_SINGULARITY = 0
TYPE_1_SINGULARITY = 1
TYPE_2_SINGULARITY = 2


def _inv(A):
    """Eigen dynamic-size .inverse() (PartialPivLU)."""
    return np.linalg.inv(np.asarray(A, dtype=np.float64))


def _bie_mass(robot: SaiModel, thr: float):
    """reference JointTask.cpp:254-259 / SingularityHandler.cpp:176-181"""
    M_BIE = robot.M()
    for i in range(robot.dof()):
        if M_BIE[i, i] < thr:
            M_BIE[i, i] = thr
    return M_BIE


class POPCExplicitForceControl:
    """reference src/helper_modules/POPCExplicitForceControl.cpp:5-96"""

    PO_window_size = 250   # POPCExplicitForceControl.h:37
    PO_max_counter = 50    # POPCExplicitForceControl.h:38

    def __init__(self, loop_timestep):
        self._loop_timestep = loop_timestep
        self._is_enabled = False
        self.reInitialize()

    def reInitialize(self):                                     # :10-21
        self._passivity_observer_value = 0.0
        self._E_correction = 0.0
        self._stored_energy_PO = 0.0
        self._PO_buffer_window = deque()
        self._PO_counter = self.PO_max_counter
        self._Rc = 1.0
        self._vcl_squared_sum = 0.0

    def enable(self):
        self._is_enabled = True

    def disable(self):
        self._is_enabled = False
        self.reInitialize()

    def computePassivitySaturatedForce(self, fd, fs, vcl, vr, kv_force, k_feedforward):   # :30-96
        if not self._is_enabled:
            return vcl - kv_force @ vr
        dt = self._loop_timestep
        F_cmd = k_feedforward * fd + self._Rc * vcl - kv_force @ vr
        vc_squared = float(vcl @ vcl)
        f_diff = fs - fd
        power_input_output = (float(f_diff @ vcl) - float(F_cmd @ vr)) * dt
        self._passivity_observer_value += power_input_output
        self._PO_buffer_window.append(power_input_output)
        if self._passivity_observer_value + self._stored_energy_PO + self._E_correction > 0:
            while len(self._PO_buffer_window) > self.PO_window_size:
                if self._passivity_observer_value + self._E_correction + self._stored_energy_PO > self._PO_buffer_window[0]:
                    if self._PO_buffer_window[0] > 0:
                        self._passivity_observer_value -= self._PO_buffer_window[0]
                    self._PO_buffer_window.popleft()
                else:
                    break
        if self._PO_counter <= 0:
            self._PO_counter = self.PO_max_counter
            old_Rc = self._Rc
            tot = self._passivity_observer_value + self._stored_energy_PO + self._E_correction
            if tot < 0:
                with np.errstate(divide="ignore", invalid="ignore"):
                    self._Rc = float(np.float64(1.0) + np.float64(tot) / np.float64(self._vcl_squared_sum * dt))
                if self._Rc > 1:
                    self._Rc = 1.0
                if self._Rc < 0:
                    self._Rc = 0.0
            else:
                self._Rc = (1 + (0.1 * self.PO_max_counter - 1) * self._Rc) / float(0.1 * self.PO_max_counter)
            self._E_correction += (1 - old_Rc) * self._vcl_squared_sum * dt
            self._vcl_squared_sum = 0.0
        self._PO_counter -= 1
        self._vcl_squared_sum += vc_squared
        return self._Rc * vcl - kv_force @ vr


class SingularityHandler:
    """reference src/tasks/SingularityHandler.cpp:24-368"""

    def __init__(self, robot: SaiModel, link_name, compliant_frame, task_rank):
        self._robot = robot
        self._link_name = link_name
        self._compliant_R, self._compliant_t = compliant_frame
        self._task_rank = task_rank
        n = self._dof = robot.dof()
        # :10-20 constants, :66-72
        self._s_abs_tol = 1e-3
        self._type_1_tol = 0.5
        # Appendix C3: the reference reads _type_2_torque_ratio before initialising it
        # (SingularityHandler.cpp:48 vs :69); the intended 1e-2 is used.
        self._type_2_torque_ratio = 1e-2
        self._type_2_angle_threshold = 5 * math.pi / 180
        self._perturb_step_size = 5.0
        self._buffer_size = 200
        self._q_upper = np.zeros(n); self._q_lower = np.zeros(n)
        self._tau_upper = np.zeros(n); self._tau_lower = np.zeros(n)
        self._joint_midrange = np.zeros(n); self._type_2_torque_vector = np.zeros(n)
        for i, lim in enumerate(robot.jointLimits()):             # :43-51
            self._q_upper[i] = lim.position_upper
            self._q_lower[i] = lim.position_lower
            self._joint_midrange[i] = 0.5 * (lim.position_lower + lim.position_upper)
            self._type_2_torque_vector[i] = self._type_2_torque_ratio * lim.effort
            self._tau_upper[i] = lim.effort
            self._tau_lower[i] = -lim.effort
        self._singularity_types = []
        self._q_prior = self._joint_midrange.copy()
        self._dq_prior = np.zeros(n)
        self.setSingularityHandlingGains(50.0, 14.0, 5.0)
        self._dynamic_decoupling_type = BOUNDED_INERTIA_ESTIMATES
        self._bie_threshold = 0.1
        self._type_1_counter = 0
        self._type_2_counter = 0
        self._type_2_direction = np.ones(n)
        self._enforce_type_1_strategy = False
        self._enforce_handling_strategy = True
        self._singularity_history = deque()
        self._s_min, self._s_max = 6e-3, 6e-2
        self._alpha = 1.0
        self._N = np.zeros((n, n))
        r = task_rank
        self._task_range_ns = np.zeros((6, 1)); self._task_range_s = np.zeros((6, 1))
        self._joint_task_range_s = np.zeros((n, 1))
        self._projected_jacobian_ns = np.zeros((1, n)); self._projected_jacobian_s = np.zeros((1, n))
        self._Lambda_ns = np.zeros((r, r)); self._Lambda_s = np.zeros((r, r))
        self._N_ns = np.eye(n)
        self._Lambda_joint_s = np.zeros((1, 1))
        self._posture_projected_jacobian = np.zeros((1, n))
        self._svd_s = np.zeros(min(6, n))

    def setDynamicDecouplingType(self, t):
        self._dynamic_decoupling_type = t

    def setBoundedInertiaEstimateThreshold(self, thr):
        self._bie_threshold = thr

    def getNullspace(self):
        return self._N

    def setSingularityHandlingBounds(self, s_min, s_max):
        self._s_min, self._s_max = s_min, s_max

    def setSingularityHandlingGains(self, kp1, kv1, kv2):
        self._kp_type_1, self._kv_type_1, self._kv_type_2 = kp1, kv1, kv2

    def handleAllSingularitiesAsType1(self, flag):
        self._enforce_type_1_strategy = flag

    def setType1Posture(self, q_des):
        self._q_prior = np.asarray(q_des, dtype=np.float64).copy()

    def enableSingularityHandling(self):
        self._enforce_handling_strategy = True

    def disableSingularityHandling(self):
        self._enforce_handling_strategy = False

    def updateTaskModel(self, projected_jacobian, N_prec):        # :75-228
        robot = self._robot
        r = self._task_rank
        n = self._dof
        U, s, Vt = np.linalg.svd(projected_jacobian, full_matrices=False)
        V = Vt.T.copy()
        U = U.copy()
        # SIGN CONVENTION (restatement decision): Eigen::JacobiSVD leaves the sign of each
        # singular-vector pair unspecified, yet classifySingularity() below perturbs q by +5 rad along V_s[:, i]
        # (:254), so the reference's type-1/type-2 decision depends on the sign its SVD happens to return.
        # Every implementation in this repo orients each pair so that the largest-magnitude entry of V[:, i]
        # is positive.
        for k in range(V.shape[1]):
            j = int(np.argmax(np.abs(V[:, k])))
            if V[j, k] < 0:
                V[:, k] = -V[:, k]
                U[:, k] = -U[:, k]
        self._svd_U, self._svd_s, self._svd_V = U, s, V
        Minv = robot.MInv()

        if s[0] < self._s_abs_tol:                                # :83-98
            self._alpha = 0.0
            self._task_range_ns = np.zeros((r, 1))
            self._projected_jacobian_ns = np.zeros((r, n))
            self._Lambda_ns = np.zeros((r, r))
            self._task_range_s = U[:, :r].copy()
            self._joint_task_range_s = V[:, :r].copy()
            self._projected_jacobian_s = self._task_range_s.T @ projected_jacobian
            self._Lambda_s = np.linalg.pinv(self._projected_jacobian_s @ Minv @ self._projected_jacobian_s.T)
        else:
            # Appendix C6: for task_rank == 1 the reference loop body never runs
            # (SingularityHandler.cpp:100); the intended "always non-singular" is used.
            rng = range(1, r) if r > 1 else [0]
            for i in rng:
                inv_condition_number = s[i] / s[0]
                if r > 1 and inv_condition_number < self._s_max:  # :103-121
                    self._alpha = min(max((inv_condition_number - self._s_min) / (self._s_max - self._s_min), 0.0), 1.0)
                    self._task_range_ns = U[:, :i].copy()
                    self._projected_jacobian_ns = self._task_range_ns.T @ projected_jacobian
                    ns = robot.operationalSpaceMatrices(self._projected_jacobian_ns)
                    self._Lambda_ns, self._Jbar_ns, self._N_ns = ns.Lambda, ns.Jbar, ns.N
                    self._task_range_s = U[:, i:r].copy()
                    self._joint_task_range_s = V[:, i:r].copy()
                    self._projected_jacobian_s = self._task_range_s.T @ projected_jacobian
                    self._Lambda_s = _inv(self._projected_jacobian_s @ Minv @ self._projected_jacobian_s.T)
                    break
                elif i == r - 1:                                  # :123-141
                    self._alpha = 1.0
                    self._task_range_ns = U[:, :r].copy()
                    self._projected_jacobian_ns = self._task_range_ns.T @ projected_jacobian
                    ns = robot.operationalSpaceMatrices(self._projected_jacobian_ns)
                    self._Lambda_ns, self._Jbar_ns, self._N_ns = ns.Lambda, ns.Jbar, ns.N
                    self._task_range_s = np.zeros((r, 1))
                    self._joint_task_range_s = np.zeros((n, 1))
                    self._projected_jacobian_s = np.zeros((r, n))
                    self._Lambda_s = np.zeros((r, r))

        # model updates :145-158
        if np.linalg.norm(self._task_range_s) == 0 or not self._enforce_handling_strategy:
            self._N = self._N_ns
            self._Lambda_joint_s = np.zeros((1, 1))
        elif np.linalg.norm(self._task_range_ns) == 0:
            self._N = N_prec
            self._Lambda_joint_s = np.zeros((1, 1))
        else:
            self._posture_projected_jacobian = self._joint_task_range_s.T @ self._N_ns @ N_prec
            osm = robot.operationalSpaceMatrices(self._posture_projected_jacobian)
            self._Lambda_joint_s = osm.Lambda
            self._N = osm.N @ self._N_ns

        dt = self._dynamic_decoupling_type                        # :160-225
        if dt == FULL_DYNAMIC_DECOUPLING:
            self._Lambda_ns_modified = self._Lambda_ns
            self._Lambda_s_modified = self._Lambda_s
            self._Lambda_joint_s_modified = self._Lambda_joint_s
        elif dt == IMPEDANCE:
            self._Lambda_ns_modified = np.eye(self._task_range_ns.shape[1])
            self._Lambda_s_modified = np.eye(self._task_range_s.shape[1])
            self._Lambda_joint_s_modified = np.eye(self._joint_task_range_s.shape[1])
        else:
            M_inv_BIE = _inv(_bie_mass(robot, self._bie_threshold))
            if np.linalg.norm(self._task_range_ns) != 0:
                self._Lambda_ns_modified = _inv(self._projected_jacobian_ns @ M_inv_BIE @ self._projected_jacobian_ns.T)
            else:
                self._Lambda_ns_modified = self._Lambda_ns
            if np.linalg.norm(self._task_range_s) != 0:
                self._Lambda_s_modified = _inv(self._projected_jacobian_s @ M_inv_BIE @ self._projected_jacobian_s.T)
            else:
                self._Lambda_s_modified = self._Lambda_s
            if np.linalg.norm(self._task_range_s) != 0 and np.linalg.norm(self._task_range_ns) != 0 \
                    and self._enforce_handling_strategy:
                # Appendix C7: the reference also evaluates this with a stale posture
                # Jacobian when the branch above did not refresh it; the value is then
                # never consumed (computeTorques returns before using it).
                self._Lambda_joint_s_modified = _inv(
                    self._posture_projected_jacobian @ M_inv_BIE @ self._posture_projected_jacobian.T)
            else:
                self._Lambda_joint_s_modified = self._Lambda_joint_s

        self.classifySingularity(self._task_range_s, self._joint_task_range_s)

    def classifySingularity(self, singular_task_range, singular_joint_task_range):   # :230-295
        robot = self._robot
        if len(self._singularity_types) == 0 or (self._type_2_counter > self._type_1_counter):
            self._q_prior = robot.q()
            self._dq_prior = robot.dq()
        if np.linalg.norm(singular_task_range) == 0:
            self._singularity_types = []
            self._singularity_history.clear()
            self._type_1_counter = 0
            self._type_2_counter = 0
            return
        k = singular_task_range.shape[1]
        self._singularity_types = [NO_SINGULARITY] * k
        curr_q = robot.q()
        curr_pos = robot.position(self._link_name, self._compliant_t)
        curr_ori = robot.rotation(self._link_name, self._compliant_R)
        for i in range(k):
            delta_q = self._perturb_step_size * singular_joint_task_range[:, i]
            robot.setQ(curr_q + delta_q)
            robot.updateKinematics()
            pos_delta = robot.position(self._link_name, self._compliant_t) - curr_pos
            ori_delta = orientationError(robot.rotation(self._link_name, self._compliant_R), curr_ori)
            delta_vector = np.concatenate([pos_delta, ori_delta])
            motion = abs(float(delta_vector @ singular_task_range[:, i]))
            self._singularity_types[i] = TYPE_1_SINGULARITY if motion > self._type_1_tol else TYPE_2_SINGULARITY
            robot.setQ(curr_q)
            robot.updateKinematics()
        if TYPE_1_SINGULARITY in self._singularity_types:
            self._singularity_history.append(TYPE_1_SINGULARITY)
            self._type_1_counter += 1
        else:
            self._singularity_history.append(TYPE_2_SINGULARITY)
            self._type_2_counter += 1
        if len(self._singularity_history) > self._buffer_size:
            if self._singularity_history[0] == TYPE_1_SINGULARITY:
                self._type_1_counter -= 1
            elif self._singularity_history[0] == TYPE_2_SINGULARITY:
                self._type_2_counter -= 1
            self._singularity_history.popleft()

    def computeTorques(self, unit_mass_force, force_related_terms):                  # :297-368
        robot = self._robot
        Jns, Uns = self._projected_jacobian_ns, self._task_range_ns
        if len(self._singularity_types) == 0:
            return Jns.T @ (self._Lambda_ns_modified @ Uns.T @ unit_mass_force + Uns.T @ force_related_terms)
        elif self._dynamic_decoupling_type == IMPEDANCE:
            return Jns.T @ (Uns.T @ unit_mass_force + Uns.T @ force_related_terms)
        tau_ns = np.zeros(self._dof)
        if np.linalg.norm(Uns) == 0:
            return tau_ns
        tau_ns = Jns.T @ (self._Lambda_ns_modified @ Uns.T @ unit_mass_force + Uns.T @ force_related_terms)
        if not self._enforce_handling_strategy:
            return tau_ns
        Vs, Us = self._joint_task_range_s, self._task_range_s
        Jpost = self._posture_projected_jacobian
        q, dq = robot.q(), robot.dq()
        if self._type_1_counter > self._type_2_counter or self._enforce_type_1_strategy:
            unit_torques = -self._kp_type_1 * (q - self._q_prior) - self._kv_type_1 * dq
            joint_strategy_torques = Jpost.T @ self._Lambda_joint_s_modified @ Vs.T @ unit_torques
        else:
            for i in range(Vs.shape[0]):
                if Vs[i, 0] != 0:
                    if abs(q[i] - self._q_upper[i]) < self._type_2_angle_threshold:
                        self._type_2_direction[i] = -1
                    elif abs(q[i] - self._q_lower[i]) < self._type_2_angle_threshold:
                        self._type_2_direction[i] = 1
            f = unit_mass_force + force_related_terms
            nf = np.linalg.norm(f)
            fTd = float((f / nf if nf > 0 else f) @ Us[:, 0])   # Eigen normalized(): zero stays zero
            magnitude_unit_torques = abs(fTd) * self._type_2_torque_vector
            unit_torques = self._type_2_direction * magnitude_unit_torques
            joint_strategy_torques = Jpost.T @ Vs.T @ unit_torques + \
                Jpost.T @ self._Lambda_joint_s_modified @ Vs.T @ (-self._kv_type_2 * dq)
        singular_task_torques = self._projected_jacobian_s.T @ (
            self._Lambda_s_modified @ Us.T @ unit_mass_force + Us.T @ force_related_terms)
        for i in range(self._dof):
            if math.isnan(singular_task_torques[i]):
                singular_task_torques[i] = 0
            elif singular_task_torques[i] > self._tau_upper[i]:
                singular_task_torques[i] = self._tau_upper[i]
            elif singular_task_torques[i] < self._tau_lower[i]:
                singular_task_torques[i] = self._tau_lower[i]
        return tau_ns + self._alpha * singular_task_torques + (1 - self._alpha) * joint_strategy_torques


class JointTask:
    """reference src/tasks/JointTask.cpp:14-356, defaults JointTask.h:31-45
    (internal OTG default differs: OFF, see module docstring)."""

    task_type = "joint"

    def __init__(self, robot: SaiModel, joint_selection_matrix=None, task_name="joint_task", loop_timestep=0.001):
        self._robot = robot
        self._task_name = task_name
        self._loop_timestep = loop_timestep
        n = robot.dof()
        if joint_selection_matrix is None:
            self._joint_selection = np.eye(n)
        else:
            S = np.atleast_2d(np.asarray(joint_selection_matrix, dtype=np.float64))
            if S.shape[1] != n:
                raise ValueError("joint selection matrix size not consistent with robot dof")
            if np.linalg.matrix_rank(S) != S.shape[0]:
                raise ValueError("joint selection matrix is not full rank")
            self._joint_selection = S
        self._task_dof = self._joint_selection.shape[0]
        self._dynamic_decoupling_type = BOUNDED_INERTIA_ESTIMATES
        self._bie_threshold = 0.1
        self.setGains(50.0, 14.0, 0.0)
        self._use_velocity_saturation_flag = False
        self._saturation_velocity = np.full(self._task_dof, math.pi / 3)
        self._N_prec = np.eye(n)
        self._M_partial = np.eye(self._task_dof)
        self._M_partial_modified = np.eye(self._task_dof)
        self._projected_jacobian = self._joint_selection.copy()
        self._N = np.zeros((n, n))
        self._current_task_range = np.eye(self._task_dof)
        self.reInitializeTask()

    def getLoopTimestep(self):
        return self._loop_timestep

    def getTaskName(self):
        return self._task_name

    def isFullJointTask(self):
        return self._task_dof == self._robot.dof()

    def reInitializeTask(self):                                   # :91-107
        self._current_position = self._joint_selection @ self._robot.q()
        self._current_velocity = np.zeros(self._task_dof)
        self._goal_position = self._current_position.copy()
        self._goal_velocity = np.zeros(self._task_dof)
        self._goal_acceleration = np.zeros(self._task_dof)
        self._integrated_position_error = np.zeros(self._task_dof)

    def setGoalPosition(self, v):
        v = np.asarray(v, dtype=np.float64)
        if v.shape != (self._task_dof,):
            raise ValueError("goal position vector size not consistent with task dof")
        self._goal_position = v.copy()

    def setGoalVelocity(self, v):
        v = np.asarray(v, dtype=np.float64)
        if v.shape != (self._task_dof,):
            raise ValueError("goal velocity vector size not consistent with task dof")
        self._goal_velocity = v.copy()

    def setGoalAcceleration(self, v):
        v = np.asarray(v, dtype=np.float64)
        if v.shape != (self._task_dof,):
            raise ValueError("goal acceleration vector size not consistent with task dof")
        self._goal_acceleration = v.copy()

    def setGains(self, kp, kv, ki=0.0):                           # :158-205
        kp, kv, ki = (np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in (kp, kv, ki))
        if kp.size == 1 and kv.size == 1 and ki.size == 1:
            if kp[0] < 0 or kv[0] < 0 or ki[0] < 0:
                raise ValueError("gains must be positive or zero")
            k = self._task_dof
            self._kp, self._kv, self._ki = kp[0] * np.eye(k), kv[0] * np.eye(k), ki[0] * np.eye(k)
            return
        if kp.size != self._task_dof or kv.size != self._task_dof or ki.size != self._task_dof:
            raise ValueError("size of gain vectors inconsistent with number of task dofs")
        if kp.max() < 0 or kv.max() < 0 or ki.max() < 0:
            raise ValueError("gains must be positive or zero")
        self._kp, self._kv, self._ki = np.diag(kp), np.diag(kv), np.diag(ki)

    def setDynamicDecouplingType(self, t):
        self._dynamic_decoupling_type = t

    def setBoundedInertiaEstimateThreshold(self, thr):            # JointTask.h:372-378
        self._bie_threshold = 0.0 if thr < 0 else thr

    def enableVelocitySaturation(self, sat):                      # :408-435
        sat = np.atleast_1d(np.asarray(sat, dtype=np.float64))
        if sat.size == 1:
            sat = np.full(self._task_dof, sat[0])
        if sat.size != self._task_dof:
            raise ValueError("saturation velocity vector size not consistent with task dof")
        if sat.min() <= 0:
            raise ValueError("saturation velocity must be positive")
        self._use_velocity_saturation_flag = True
        self._saturation_velocity = sat.copy()

    def disableVelocitySaturation(self):
        self._use_velocity_saturation_flag = False

    def getTaskNullspace(self):
        return self._N

    def getPreviousTasksNullspace(self):                          # JointTask.h:223
        return self._N_prec

    def getTaskAndPreviousNullspace(self):                        # JointTask.h:224-226
        return self._N @ self._N_prec

    # JointTask.h:108-190
    def getCurrentPosition(self): return self._current_position
    def getCurrentVelocity(self): return self._current_velocity
    def getGoalPosition(self): return self._goal_position

    def updateTaskModel(self, N_prec):                            # :218-283
        robot = self._robot
        n = robot.dof()
        N_prec = np.asarray(N_prec, dtype=np.float64)
        if N_prec.shape[0] != N_prec.shape[1]:
            raise ValueError("N_prec matrix not square")
        if N_prec.shape[0] != n:
            raise ValueError("N_prec matrix size not consistent with robot dof")
        self._N_prec = N_prec.copy()
        self._projected_jacobian = self._joint_selection @ self._N_prec
        self._current_task_range = matrixRangeBasis(self._projected_jacobian)
        if np.linalg.norm(self._current_task_range) == 0:
            self._N = np.eye(n)
            return
        U = self._current_task_range
        osm = robot.operationalSpaceMatrices(U.T @ self._projected_jacobian)
        self._M_partial = osm.Lambda
        self._N = osm.N
        if self._dynamic_decoupling_type == FULL_DYNAMIC_DECOUPLING:
            self._M_partial_modified = self._M_partial
        elif self._dynamic_decoupling_type == BOUNDED_INERTIA_ESTIMATES:
            M_inv_BIE = _inv(_bie_mass(robot, self._bie_threshold))
            self._M_partial_modified = _inv(U.T @ self._projected_jacobian @ M_inv_BIE @ self._projected_jacobian.T @ U)
        else:
            self._M_partial_modified = np.eye(U.shape[1])

    def computeTorques(self, tau_prec=None):
        if tau_prec is not None:                                  # :285-292
            task_torques = self.computeTorques()
            U = self._current_task_range
            if np.linalg.norm(U) == 0:
                # the reference would multiply by a stale/placeholder range here; with a
                # zero range the product is zero.
                return task_torques
            comp = self._projected_jacobian.T @ U @ self._M_partial @ U.T @ self._joint_selection @ \
                self._robot.MInv() @ np.asarray(tau_prec, dtype=np.float64)
            return task_torques - comp
        robot = self._robot                                       # :294-356
        self._projected_jacobian = self._joint_selection @ self._N_prec
        self._current_position = self._joint_selection @ robot.q()
        self._current_velocity = self._joint_selection @ robot.dq()
        U = self._current_task_range
        if np.linalg.norm(U) == 0:
            return np.zeros(robot.dof())
        desired_position = self._goal_position
        desired_velocity = self._goal_velocity
        desired_acceleration = self._goal_acceleration
        self._integrated_position_error = self._integrated_position_error + \
            (self._current_position - desired_position) * self._loop_timestep
        if self._use_velocity_saturation_flag:
            kv_inverse = computePseudoInverse(self._kv)
            desired_velocity = -self._kp @ kv_inverse @ (self._current_position - desired_position) - \
                self._ki @ kv_inverse @ self._integrated_position_error
            # Appendix C8: the reference loops to robot dof (JointTask.cpp:332); task dof is used.
            for i in range(self._task_dof):
                if desired_velocity[i] > self._saturation_velocity[i]:
                    desired_velocity[i] = self._saturation_velocity[i]
                elif desired_velocity[i] < -self._saturation_velocity[i]:
                    desired_velocity[i] = -self._saturation_velocity[i]
            t = -self._kv @ (self._current_velocity - desired_velocity)
        else:
            t = -self._kp @ (self._current_position - desired_position) - \
                self._kv @ (self._current_velocity - desired_velocity) - \
                self._ki @ self._integrated_position_error
        self._pid_torques = t                                     # kept for tests
        f = self._M_partial @ U.T @ desired_acceleration + self._M_partial_modified @ U.T @ t
        return self._projected_jacobian.T @ U @ f


class MotionForceTask:
    """reference src/tasks/MotionForceTask.cpp:16-1001, defaults MotionForceTask.h:40-75
    (internal OTG default differs: OFF)."""

    task_type = "motion_force"

    def __init__(self, robot: SaiModel, link_name, compliant_frame=None,
                 controlled_directions_translation=None, controlled_directions_rotation=None,
                 task_name="motion_force_task", is_force_motion_parametrization_in_compliant_frame=False,
                 loop_timestep=0.001):
        self._robot = robot
        self._link_name = link_name
        self._task_name = task_name
        self._loop_timestep = loop_timestep
        if compliant_frame is None:
            compliant_frame = (np.eye(3), np.zeros(3))
        self._compliant_R = np.asarray(compliant_frame[0], dtype=np.float64)
        self._compliant_t = np.asarray(compliant_frame[1], dtype=np.float64)
        self._in_compliant = bool(is_force_motion_parametrization_in_compliant_frame)
        if controlled_directions_translation is None and controlled_directions_rotation is None:
            self._partial_task_projection = np.eye(6)                           # :28
        else:
            dt_ = list(controlled_directions_translation or [])
            dr_ = list(controlled_directions_rotation or [])
            if len(dt_) == 0 and len(dr_) == 0:
                raise ValueError("controlled directions cannot both be empty")   # :47-53
            bt = np.zeros((3, 1)); br = np.zeros((3, 1))
            if dt_:
                bt = matrixRangeBasis(np.array(dt_, dtype=np.float64).T)
            if dr_:
                br = matrixRangeBasis(np.array(dr_, dtype=np.float64).T)
            P = np.zeros((6, 6))
            P[:3, :3] = bt @ bt.T
            P[3:, 3:] = br @ br.T
            self._partial_task_projection = P
        self._initialSetup()

    def _initialSetup(self):                                      # :92-202
        robot = self._robot
        n = robot.dof()
        self._T_cs_R, self._T_cs_t = np.eye(3), np.zeros(3)
        self._POPC_force = POPCExplicitForceControl(self._loop_timestep)
        self.setPosControlGains(100.0, 20.0, 0.0)
        self.setOriControlGains(200.0, 28.3, 0.0)
        self.setForceControlGains(0.7, 10.0, 1.3)
        self.setMomentControlGains(0.7, 10.0, 1.3)
        self._use_velocity_saturation_flag = False
        self._linear_saturation_velocity = 0.3
        self._angular_saturation_velocity = math.pi / 3
        self._kff_force = 0.95
        self._kff_moment = 0.95
        self._max_force_control_feedback_output = 20.0
        self._max_moment_control_feedback_output = 10.0
        self._force_space_dimension = 0
        self._moment_space_dimension = 0
        self._force_or_motion_axis = np.zeros(3)
        self._moment_or_rotmotion_axis = np.zeros(3)
        self._closed_loop_force_control = False
        self._closed_loop_moment_control = False
        self._jacobian = np.zeros((6, n))
        self._projected_jacobian = np.zeros((6, n))
        self._Lambda = np.zeros((6, 6))        # never written again (Appendix C1)
        self._N = np.zeros((n, n))
        self._N_prec = np.eye(n)
        range_pos = matrixRangeBasis(self._partial_task_projection[:3, :3])
        range_ori = matrixRangeBasis(self._partial_task_projection[3:, 3:])
        self._pos_range = 0 if np.linalg.norm(range_pos) == 0 else range_pos.shape[1]
        self._ori_range = 0 if np.linalg.norm(range_ori) == 0 else range_ori.shape[1]
        if self._pos_range + self._ori_range == 0:
            raise ValueError("controlled directions cannot both be empty")
        self._singularity_handler = SingularityHandler(
            robot, self._link_name, (self._compliant_R, self._compliant_t), self._pos_range + self._ori_range)
        self.setSingularityHandlingBounds(6e-3, 6e-2)
        self.setDynamicDecouplingType(BOUNDED_INERTIA_ESTIMATES)
        self.setBoundedInertiaEstimateThreshold(0.1)
        self._integrated_force_error = np.zeros(3)
        self._integrated_moment_error = np.zeros(3)
        self.reInitializeTask()

    def getLoopTimestep(self):
        return self._loop_timestep

    def getTaskName(self):
        return self._task_name

    def reInitializeTask(self):                                   # :204-245
        robot = self._robot
        self._current_position = robot.positionInWorld(self._link_name, self._compliant_t)
        self._goal_position = self._current_position.copy()
        self._current_orientation = robot.rotationInWorld(self._link_name, self._compliant_R)
        self._goal_orientation = self._current_orientation.copy()
        self._current_linear_velocity = np.zeros(3)
        self._goal_linear_velocity = np.zeros(3)
        self._current_angular_velocity = np.zeros(3)
        self._goal_angular_velocity = np.zeros(3)
        self._goal_linear_acceleration = np.zeros(3)
        self._goal_angular_acceleration = np.zeros(3)
        self._orientation_error = np.zeros(3)
        self._integrated_position_error = np.zeros(3)
        self._integrated_orientation_error = np.zeros(3)
        self._goal_force = np.zeros(3)
        self._sensed_force_control_world_frame = np.zeros(3)
        self._sensed_force_sensor_frame = np.zeros(3)
        self._goal_moment = np.zeros(3)
        self._sensed_moment_control_world_frame = np.zeros(3)
        self._sensed_moment_sensor_frame = np.zeros(3)
        self.resetIntegrators()
        self._unit_mass_force = np.zeros(6)

    # ---- setters mirrored from MotionForceTask.h:211-247 ----
    def setGoalPosition(self, v): self._goal_position = np.asarray(v, dtype=np.float64).copy()
    def setGoalOrientation(self, R): self._goal_orientation = np.asarray(R, dtype=np.float64).copy()
    def setGoalLinearVelocity(self, v): self._goal_linear_velocity = np.asarray(v, dtype=np.float64).copy()
    def setGoalAngularVelocity(self, v): self._goal_angular_velocity = np.asarray(v, dtype=np.float64).copy()
    def setGoalLinearAcceleration(self, v): self._goal_linear_acceleration = np.asarray(v, dtype=np.float64).copy()
    def setGoalAngularAcceleration(self, v): self._goal_angular_acceleration = np.asarray(v, dtype=np.float64).copy()
    def setGoalForce(self, v): self._goal_force = np.asarray(v, dtype=np.float64).copy()
    def setGoalMoment(self, v): self._goal_moment = np.asarray(v, dtype=np.float64).copy()

    @staticmethod
    def _gain3(kp, kv, ki, what):
        kp, kv, ki = (np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in (kp, kv, ki))
        if kp.size == 1 and kv.size == 1 and ki.size == 1:
            if kp[0] < 0 or kv[0] < 0 or ki[0] < 0:
                raise ValueError("all gains should be positive or zero in " + what)
            return kp[0] * np.eye(3), kv[0] * np.eye(3), ki[0] * np.eye(3)
        if kp.size != 3 or kv.size != 3 or ki.size != 3:
            raise ValueError("gains should be of size 1 or 3 in " + what)
        if kp.min() < 0 or kv.min() < 0 or ki.min() < 0:
            raise ValueError("all gains should be positive or zero in " + what)
        return np.diag(kp), np.diag(kv), np.diag(ki)

    def setPosControlGains(self, kp, kv, ki=0.0):                 # :581-628
        self._kp_pos, self._kv_pos, self._ki_pos = self._gain3(kp, kv, ki, "setPosControlGains")

    def setOriControlGains(self, kp, kv, ki=0.0):                 # :668-715
        self._kp_ori, self._kv_ori, self._ki_ori = self._gain3(kp, kv, ki, "setOriControlGains")

    def setForceControlGains(self, kp, kv, ki):                   # MotionForceTask.h:305-310
        self._kp_force, self._kv_force, self._ki_force = kp * np.eye(3), kv * np.eye(3), ki * np.eye(3)

    def setMomentControlGains(self, kp, kv, ki):                  # MotionForceTask.h:319-324
        self._kp_moment, self._kv_moment, self._ki_moment = kp * np.eye(3), kv * np.eye(3), ki * np.eye(3)

    def setFeedforwardForceGain(self, k): self._kff_force = k
    def setFeedforwardmomentGain(self, k): self._kff_moment = k
    def setMaxForceControlFeedbackOutput(self, v): self._max_force_control_feedback_output = v
    def setMaxMomentControlFeedbackOutput(self, v): self._max_moment_control_feedback_output = v

    def enableVelocitySaturation(self, linear_vel_sat=0.3, angular_vel_sat=math.pi / 3):   # :771-792
        if linear_vel_sat <= 0 or angular_vel_sat <= 0:
            raise ValueError("Velocity saturation values should be strictly positive")
        self._use_velocity_saturation_flag = True
        self._linear_saturation_velocity = linear_vel_sat
        self._angular_saturation_velocity = angular_vel_sat

    def disableVelocitySaturation(self): self._use_velocity_saturation_flag = False
    def enablePassivity(self): self._POPC_force.enable()
    def disablePassivity(self): self._POPC_force.disable()
    def setDynamicDecouplingType(self, t): self._singularity_handler.setDynamicDecouplingType(t)
    def setBoundedInertiaEstimateThreshold(self, t): self._singularity_handler.setBoundedInertiaEstimateThreshold(t)
    def handleAllSingularitiesAsType1(self, f): self._singularity_handler.handleAllSingularitiesAsType1(f)
    def setType1Posture(self, q): self._singularity_handler.setType1Posture(q)
    def enableSingularityHandling(self): self._singularity_handler.enableSingularityHandling()
    def disableSingularityHandling(self): self._singularity_handler.disableSingularityHandling()
    def setSingularityHandlingBounds(self, a, b): self._singularity_handler.setSingularityHandlingBounds(a, b)
    def setSingularityHandlingGains(self, a, b, c): self._singularity_handler.setSingularityHandlingGains(a, b, c)

    def setForceSensorFrame(self, link_name, T_in_link):          # :794-803
        if link_name != self._link_name:
            raise ValueError("sensor link must be the control link")
        R_l, t_l = np.asarray(T_in_link[0], dtype=np.float64), np.asarray(T_in_link[1], dtype=np.float64)
        # compliant_frame.inverse() * T
        self._T_cs_R = self._compliant_R.T @ R_l
        self._T_cs_t = self._compliant_R.T @ (t_l - self._compliant_t)

    def updateSensedForceAndMoment(self, f_sensor, m_sensor):     # :805-828
        f_sensor = np.asarray(f_sensor, dtype=np.float64); m_sensor = np.asarray(m_sensor, dtype=np.float64)
        self._sensed_force_sensor_frame = f_sensor.copy()
        self._sensed_moment_sensor_frame = m_sensor.copy()
        R_wl, _ = self._robot.transformInWorld(self._link_name)
        R_wc = R_wl @ self._compliant_R
        f = self._T_cs_R @ f_sensor
        m = np.cross(self._T_cs_t, f) + self._T_cs_R @ m_sensor
        self._sensed_force_control_world_frame = R_wc @ f
        self._sensed_moment_control_world_frame = R_wc @ m

    def parametrizeForceMotionSpaces(self, dim, axis=(0.0, 0.0, 0.0)):   # :830-858
        if dim < 0 or dim > 3:
            raise ValueError("Force space dimension should be between 0 and 3")
        reset = dim != self._force_space_dimension
        self._force_space_dimension = dim
        if dim in (1, 2):
            axis = np.asarray(axis, dtype=np.float64)
            if np.linalg.norm(axis) < 1e-2:
                raise ValueError("Force or motion axis should be a non singular vector")
            a = axis / np.linalg.norm(axis)
            reset = reset or not np.allclose(a, self._force_or_motion_axis, rtol=1e-12, atol=0)
            self._force_or_motion_axis = a
        if reset:
            self._goal_position = self._current_position.copy()
            self._goal_linear_velocity = np.zeros(3)
            self._goal_linear_acceleration = np.zeros(3)
            self.resetIntegratorsLinear()
        return reset

    def parametrizeMomentRotMotionSpaces(self, dim, axis=(0.0, 0.0, 0.0)):   # :860-890
        if dim < 0 or dim > 3:
            raise ValueError("Moment space dimension should be between 0 and 3")
        reset = dim != self._moment_space_dimension
        self._moment_space_dimension = dim
        if dim in (1, 2):
            axis = np.asarray(axis, dtype=np.float64)
            if np.linalg.norm(axis) < 1e-2:
                raise ValueError("Moment or rot motion axis should be a non singular vector")
            a = axis / np.linalg.norm(axis)
            reset = reset or not np.allclose(a, self._moment_or_rotmotion_axis, rtol=1e-12, atol=0)
            self._moment_or_rotmotion_axis = a
        if reset:
            self._goal_orientation = self._current_orientation.copy()
            self._goal_angular_velocity = np.zeros(3)
            self._goal_angular_acceleration = np.zeros(3)
            self.resetIntegratorsAngular()
        return reset

    def setClosedLoopForceControl(self, flag=True):               # :973-979
        if self._closed_loop_force_control != flag:
            self._closed_loop_force_control = flag
            self.resetIntegratorsLinear()

    def setClosedLoopMomentControl(self, flag=True):              # :980-986
        if self._closed_loop_moment_control != flag:
            self._closed_loop_moment_control = flag
            self.resetIntegratorsAngular()

    def resetIntegrators(self):
        self.resetIntegratorsLinear(); self.resetIntegratorsAngular()

    def resetIntegratorsLinear(self):
        self._integrated_position_error = np.zeros(3)
        self._integrated_force_error = np.zeros(3)

    def resetIntegratorsAngular(self):
        self._integrated_orientation_error = np.zeros(3)
        self._integrated_moment_error = np.zeros(3)

    def posSelectionProjector(self): return self._partial_task_projection[:3, :3]
    def oriSelectionProjector(self): return self._partial_task_projection[3:, 3:]

    def _param_rotation(self):
        if self._in_compliant:
            return self._robot.rotationInWorld(self._link_name, self._compliant_R)
        return np.eye(3)

    def getGoalForce(self): return self._param_rotation() @ self._goal_force       # :755-761
    def getGoalMoment(self): return self._param_rotation() @ self._goal_moment     # :763-769

    @staticmethod
    def _sigma(dim, P, R, a):                                     # :892-925 / :932-966
        if dim == 0:
            return np.zeros((3, 3))
        if dim == 1:
            return P @ R @ np.outer(a, a) @ R.T @ P.T
        if dim == 2:
            return P @ (np.eye(3) - R @ np.outer(a, a) @ R.T) @ P.T
        return P.copy()

    def sigmaForce(self):
        return self._sigma(self._force_space_dimension, self.posSelectionProjector(), self._param_rotation(), self._force_or_motion_axis)

    def sigmaPosition(self):                                      # :927-930
        P = self.posSelectionProjector()
        return P @ (np.eye(3) - self.sigmaForce()) @ P.T

    def sigmaMoment(self):
        return self._sigma(self._moment_space_dimension, self.oriSelectionProjector(), self._param_rotation(), self._moment_or_rotmotion_axis)

    def sigmaOrientation(self):                                   # :968-971
        P = self.oriSelectionProjector()
        return P @ (np.eye(3) - self.sigmaMoment()) @ P.T

    def getTaskNullspace(self): return self._N
    def getPreviousTasksNullspace(self): return self._N_prec                  # MotionForceTask.h:206
    def getTaskAndPreviousNullspace(self): return self._N @ self._N_prec      # MotionForceTask.h:207-209
    def getUnitMassForce(self): return self._unit_mass_force
    # MotionForceTask.h:116-140, :540-546
    def getCurrentPosition(self): return self._current_position
    def getCurrentOrientation(self): return self._current_orientation
    def getCurrentLinearVelocity(self): return self._current_linear_velocity
    def getCurrentAngularVelocity(self): return self._current_angular_velocity
    def getPositionError(self): return self.sigmaPosition() @ (self._goal_position - self._current_position)
    def getOrientationError(self): return self.sigmaOrientation() @ self._orientation_error

    def updateTaskModel(self, N_prec):                            # :247-268
        robot = self._robot
        n = robot.dof()
        N_prec = np.asarray(N_prec, dtype=np.float64)
        if N_prec.shape[0] != N_prec.shape[1]:
            raise ValueError("N_prec matrix not square")
        if N_prec.shape[0] != n:
            raise ValueError("N_prec matrix size not consistent with robot dof")
        self._N_prec = N_prec.copy()
        self._jacobian = self._partial_task_projection @ robot.JWorldFrame(self._link_name, self._compliant_t)
        self._projected_jacobian = self._jacobian @ self._N_prec
        self._singularity_handler.updateTaskModel(self._projected_jacobian, self._N_prec)
        self._N = self._singularity_handler.getNullspace()

    def computeTorques(self, tau_prec=None):
        if tau_prec is not None:                                  # :270-276
            task_torques = self.computeTorques()
            comp = self._projected_jacobian.T @ self._Lambda @ self._jacobian @ self._robot.MInv() @ \
                np.asarray(tau_prec, dtype=np.float64)            # _Lambda == 0 (Appendix C1)
            return task_torques - comp
        robot = self._robot                                       # :278-509
        dt = self._loop_timestep
        self._jacobian = self._partial_task_projection @ robot.JWorldFrame(self._link_name, self._compliant_t)
        self._projected_jacobian = self._jacobian @ self._N_prec
        self._current_position = robot.positionInWorld(self._link_name, self._compliant_t)
        self._current_orientation = robot.rotationInWorld(self._link_name, self._compliant_R)
        self._orientation_error = orientationError(self._goal_orientation, self._current_orientation)
        self._current_linear_velocity = self._jacobian[:3] @ robot.dq()
        self._current_angular_velocity = self._jacobian[3:] @ robot.dq()
        if self._pos_range + self._ori_range == 0:
            return np.zeros(robot.dof())
        sigma_force = self.sigmaForce(); sigma_moment = self.sigmaMoment()
        sigma_position = self.sigmaPosition(); sigma_orientation = self.sigmaOrientation()
        goal_force = self.getGoalForce(); goal_moment = self.getGoalMoment()

        if self._closed_loop_force_control:                       # :327-349
            self._integrated_force_error = self._integrated_force_error + \
                sigma_force @ (self._sensed_force_control_world_frame - goal_force) * dt
            fb = sigma_force @ (-self._kp_force @ (self._sensed_force_control_world_frame - goal_force)
                                - self._ki_force @ self._integrated_force_error)
            if np.linalg.norm(fb) > self._max_force_control_feedback_output:
                fb = fb * (self._max_force_control_feedback_output / np.linalg.norm(fb))
            force_feedback_related_force = self._POPC_force.computePassivitySaturatedForce(
                sigma_force @ goal_force, sigma_force @ self._sensed_force_control_world_frame,
                sigma_force @ fb, sigma_force @ self._current_linear_velocity, self._kv_force, self._kff_force)
        else:                                                     # :350-354
            force_feedback_related_force = sigma_force @ (-self._kv_force @ self._current_linear_velocity)

        if self._closed_loop_moment_control:                      # :357-378
            self._integrated_moment_error = self._integrated_moment_error + \
                sigma_moment @ (self._sensed_moment_control_world_frame - goal_moment) * dt
            mb = sigma_moment @ (-self._kp_moment @ (self._sensed_moment_control_world_frame - goal_moment)
                                 - self._ki_moment @ self._integrated_moment_error)
            if np.linalg.norm(mb) > self._max_moment_control_feedback_output:
                mb = mb * (self._max_moment_control_feedback_output / np.linalg.norm(mb))
            moment_feedback_related_force = sigma_moment @ (mb - self._kv_moment @ self._current_angular_velocity)
        else:                                                     # :379-383
            moment_feedback_related_force = sigma_moment @ (-self._kv_moment @ self._current_angular_velocity)

        desired_position = self._goal_position                    # :387-392 (OTG off)
        desired_orientation = self._goal_orientation
        desired_linear_velocity = self._goal_linear_velocity
        desired_angular_velocity = self._goal_angular_velocity
        desired_linear_acceleration = self._goal_linear_acceleration
        desired_angular_acceleration = self._goal_angular_acceleration

        self._integrated_position_error = self._integrated_position_error + \
            sigma_position @ (self._current_position - desired_position) * dt        # :411-413
        if self._use_velocity_saturation_flag:                    # :416-429
            kv_pos_inv = computePseudoInverse(self._kv_pos)
            desired_linear_velocity = -self._kp_pos @ kv_pos_inv @ sigma_position @ (self._current_position - desired_position) \
                - self._ki_pos @ kv_pos_inv @ self._integrated_position_error
            nv = np.linalg.norm(desired_linear_velocity)
            if nv > self._linear_saturation_velocity:
                desired_linear_velocity = desired_linear_velocity * (self._linear_saturation_velocity / nv)
            position_related_force = sigma_position @ (
                desired_linear_acceleration - self._kv_pos @ (self._current_linear_velocity - desired_linear_velocity))
        else:                                                     # :430-437
            position_related_force = sigma_position @ (
                desired_linear_acceleration - self._kp_pos @ (self._current_position - desired_position)
                - self._kv_pos @ (self._current_linear_velocity - desired_linear_velocity)
                - self._ki_pos @ self._integrated_position_error)

        step_orientation_error = sigma_orientation @ orientationError(desired_orientation, self._current_orientation)   # :441-443
        self._integrated_orientation_error = self._integrated_orientation_error + step_orientation_error * dt
        if self._use_velocity_saturation_flag:                    # :449-461
            kv_ori_inv = computePseudoInverse(self._kv_ori)
            desired_angular_velocity = -self._kp_ori @ kv_ori_inv @ step_orientation_error \
                - self._ki_ori @ kv_ori_inv @ self._integrated_orientation_error
            nw = np.linalg.norm(desired_angular_velocity)
            if nw > self._angular_saturation_velocity:
                desired_angular_velocity = desired_angular_velocity * (self._angular_saturation_velocity / nw)
            orientation_related_force = sigma_orientation @ (
                desired_angular_acceleration - self._kv_ori @ (self._current_angular_velocity - desired_angular_velocity))
        else:                                                     # :462-468
            orientation_related_force = sigma_orientation @ (
                desired_angular_acceleration - self._kp_ori @ step_orientation_error
                - self._kv_ori @ (self._current_angular_velocity - desired_angular_velocity)
                - self._ki_ori @ self._integrated_orientation_error)

        force_moment_contribution = np.concatenate([force_feedback_related_force, moment_feedback_related_force])
        self._unit_mass_force = np.concatenate([position_related_force, orientation_related_force])
        feedforward = np.concatenate([sigma_force @ goal_force, sigma_moment @ goal_moment])
        if self._closed_loop_force_control:                       # :484-487 (Appendix C9)
            feedforward[:3] *= self._kff_force
            feedforward[3:] *= self._kff_moment
        self._force_related_terms = force_moment_contribution + feedforward   # kept for tests
        return self._singularity_handler.computeTorques(self._unit_mass_force, self._force_related_terms)


class JointLimitAvoidanceTask:
    """reference src/tasks/JointLimitAvoidanceTask.cpp:13-421: the constraint task that
    RobotController owns.  Defaults: JointLimitAvoidanceTask.h:26-35.

    Restatement decision: `_current_task_range = matrixRangeBasis(S N_prec)` with N_prec = I is the range basis of a
    selection matrix, whose rows are orthonormal; Eigen's JacobiSVD performs no rotation on it, so the basis is the
    identity (any other orthogonal basis would silently permute the avoidance torques in :417-419).  The identity is
    used here and in the CUDA path."""

    OFF, POS_Z1, POS_Z2, VEL_Z1, VEL_Z2 = range(5)
    POSITIVE, NEGATIVE = 0, 1

    def __init__(self, robot, task_name="joint_limit_avoidance_task"):
        self._robot = robot
        self._task_name = task_name
        n = robot.dof()
        self._enabled = True
        self._kv = 20.0
        self._position_z1_to_limit = 9 * math.pi / 180.0
        self._position_z2_to_limit = 6 * math.pi / 180.0
        self._velocity_z1_to_limit = 0.5
        self._velocity_z2_to_limit = 0.3
        self._max_torque_ratio_pos_limit = 1.0
        self._max_torque_ratio_vel_limit = 0.05
        self._limit_status = [self.OFF] * n
        self._limit_direction = [self.POSITIVE] * n
        self._limit_value = [0.0] * n
        self._torque_limit_value = [0.0] * n
        self._pos_valid = {}
        self._vel_valid = {}
        for lim in robot.jointLimits():                                   # verifyValidityPerJoint :95-118
            self._pos_valid[lim.joint_index] = (lim.position_upper - lim.position_lower) > 2 * self._position_z1_to_limit
            self._vel_valid[lim.joint_index] = lim.velocity > 2 * self._velocity_z1_to_limit
        self._N_prec = np.eye(n)
        self._N = np.zeros((n, n))
        self._active_constraints = 0
        self._joint_selection = np.zeros((0, n))
        self.reInitializeTask()

    def reInitializeTask(self):                                           # :120-122
        self._computeJointSelectionMatrix()

    def _updateLimitStatus(self):                                         # :174-243
        fmax = np.finfo(np.float64).max
        for lim in self._robot.jointLimits():
            i = lim.joint_index
            self._limit_status[i] = self.OFF
            self._limit_direction[i] = self.POSITIVE
            self._limit_value[i] = 0.0
            self._torque_limit_value[i] = 0.0
            q = self._robot.q()[i]
            dq = self._robot.dq()[i]
            pv, vv = self._pos_valid[i], self._vel_valid[i]
            if pv and lim.position_upper != fmax:
                if q > lim.position_upper - self._position_z1_to_limit:
                    self._limit_direction[i] = self.POSITIVE
                    self._limit_value[i] = lim.position_upper
                    self._torque_limit_value[i] = lim.effort
                    self._limit_status[i] = self.POS_Z1
                if q > lim.position_upper - self._position_z2_to_limit:
                    self._limit_status[i] = self.POS_Z2
            if pv and lim.position_lower != -fmax:
                if q < lim.position_lower + self._position_z1_to_limit:
                    self._limit_direction[i] = self.NEGATIVE
                    self._limit_value[i] = lim.position_lower
                    self._torque_limit_value[i] = lim.effort
                    self._limit_status[i] = self.POS_Z1
                if q < lim.position_lower + self._position_z2_to_limit:
                    self._limit_status[i] = self.POS_Z2
            if vv and (self._limit_status[i] == self.OFF or self._limit_direction[i] == self.NEGATIVE):
                if dq > lim.velocity - self._velocity_z1_to_limit:
                    self._limit_direction[i] = self.POSITIVE
                    self._limit_value[i] = lim.velocity
                    self._torque_limit_value[i] = lim.effort
                    self._limit_status[i] = self.VEL_Z1
                if dq > lim.velocity - self._velocity_z2_to_limit:
                    self._limit_status[i] = self.VEL_Z2
            if vv and (self._limit_status[i] == self.OFF or self._limit_direction[i] == self.POSITIVE):
                if dq < -lim.velocity + self._velocity_z1_to_limit:
                    self._limit_direction[i] = self.NEGATIVE
                    self._limit_value[i] = -lim.velocity
                    self._torque_limit_value[i] = lim.effort
                    self._limit_status[i] = self.VEL_Z1
                if dq < -lim.velocity + self._velocity_z2_to_limit:
                    self._limit_status[i] = self.VEL_Z2
        self._active_constraints = sum(1 for s_ in self._limit_status if s_ != self.OFF)

    def _computeJointSelectionMatrix(self):                               # :245-256
        self._updateLimitStatus()
        n = self._robot.dof()
        S = np.zeros((self._active_constraints, n))
        row = 0
        for i in range(n):
            if self._limit_status[i] != self.OFF:
                S[row, i] = 1.0
                row += 1
        self._joint_selection = S

    def updateTaskModel(self, N_prec):                                    # :124-172
        n = self._robot.dof()
        N_prec = np.asarray(N_prec, dtype=np.float64)
        if N_prec.shape != (n, n):
            raise ValueError("N_prec matrix size not consistent with robot dof in JointLimitAvoidanceTask::updateTaskModel")
        if not self._enabled:
            self._N = np.eye(n)
            self._N_prec = np.eye(n)
            return
        self._N_prec = N_prec.copy()
        self._computeJointSelectionMatrix()
        self._projected_jacobian = self._joint_selection @ self._N_prec
        if self._active_constraints == 0:                                 # range basis of an empty matrix: norm 0 (:161-166)
            self._N = np.eye(n)
            return
        ops = self._robot.operationalSpaceMatrices(self._projected_jacobian)   # range basis = identity, see the class docstring
        self._M_partial = ops.Lambda
        self._N = ops.N

    def getTaskAndPreviousNullspace(self):
        return self._N @ self._N_prec

    @staticmethod
    def _blend(z, z1, z2, direction):                                     # computeBlendingCoefficient :16-37
        if direction == JointLimitAvoidanceTask.NEGATIVE:
            if z >= z1:
                return 0.0
            if z <= z2:
                return 1.0
            return (z1 - z) / (z1 - z2)
        if z <= z1:
            return 0.0
        if z >= z2:
            return 1.0
        return (z - z1) / (z2 - z1)

    def computeTorques(self, tau_tasks=None):                             # :258-421
        n = self._robot.dof()
        if tau_tasks is None:
            tau_tasks = np.zeros(n)
        if not self._enabled or self._active_constraints == 0:
            return np.zeros(n)
        q, dq = self._robot.q(), self._robot.dq()
        kv = self._kv
        pz1, pz2, vz1, vz2 = self._position_z1_to_limit, self._position_z2_to_limit, self._velocity_z1_to_limit, self._velocity_z2_to_limit
        rp, rv = self._max_torque_ratio_pos_limit, self._max_torque_ratio_vel_limit
        out = np.zeros(self._active_constraints)
        c = 0
        for i in range(n):
            st, d, lv, tl = self._limit_status[i], self._limit_direction[i], self._limit_value[i], self._torque_limit_value[i]
            clampv = lambda x: max(min(x, tl * rv), -tl * rv)
            if d == self.POSITIVE:
                if st == self.POS_Z1:
                    a = self._blend(q[i], lv - pz1, lv - pz2, d)
                    t1 = tau_tasks[i] - kv * dq[i]
                    out[c] = (1 - a) * tau_tasks[i] + a * t1
                elif st == self.POS_Z2:
                    a = self._blend(q[i], lv - pz2, lv, d)
                    t1 = tau_tasks[i] - kv * dq[i]
                    t2 = -tl * rp - kv * dq[i]
                    out[c] = (1 - a) * t1 + a * t2
                elif st == self.VEL_Z1:
                    a = self._blend(dq[i], lv - vz1, lv - vz2, d)
                    t1 = -kv * dq[i]
                    out[c] = (1 - a) * tau_tasks[i] + a * t1
                elif st == self.VEL_Z2:
                    a = self._blend(dq[i], lv - vz2, lv, d)
                    t1 = clampv(-kv * dq[i])
                    t2 = -a * tl * rv
                    out[c] = (1 - a) * t1 + a * t2
            else:
                if st == self.POS_Z1:
                    a = self._blend(q[i], lv + pz1, lv + pz2, d)
                    t1 = clampv(tau_tasks[i] - kv * dq[i])
                    out[c] = a * tau_tasks[i] + (1 - a) * t1                 # sic (:344-345)
                elif st == self.POS_Z2:
                    a = self._blend(q[i], lv + pz2, lv, d)
                    t1 = tau_tasks[i] - kv * dq[i]
                    t2 = tl * rp - kv * dq[i]
                    out[c] = (1 - a) * t1 + a * t2
                elif st == self.VEL_Z1:
                    a = self._blend(dq[i], lv + vz1, lv + vz2, d)
                    t1 = clampv(-kv * dq[i])
                    out[c] = (1 - a) * tau_tasks[i] + a * t1
                elif st == self.VEL_Z2:
                    a = self._blend(dq[i], lv + vz2, lv, d)
                    t1 = clampv(-kv * dq[i])
                    t2 = tl * rv
                    out[c] = (1 - a) * t1 + a * t2
            if st != self.OFF:
                c += 1
        return self._projected_jacobian.T @ out                           # range basis = identity


class RobotController:
    """reference src/RobotController.cpp:8-118, including the joint-limit avoidance blend (:96-112)."""

    def __init__(self, robot: SaiModel, tasks):
        if len(tasks) == 0:
            raise ValueError("RobotController must have at least one task")
        self._robot = robot
        self._enable_gravity_compensation = False
        self._enable_joint_limit_avoidance = False
        self._enable_torque_saturation = False
        self._tasks = []
        names = []
        cannot_accept_new_tasks = False
        for task in tasks:                                        # :27-59
            if task._robot is not robot:
                raise ValueError("All tasks must have the same robot model in RobotController")
            if task.getLoopTimestep() != tasks[0].getLoopTimestep():
                raise ValueError("All tasks must have the same loop timestep in RobotController")
            if task.getTaskName() in names:
                raise ValueError("Tasks in RobotController must have unique names")
            names.append(task.getTaskName())
            self._tasks.append(task)
            if cannot_accept_new_tasks:
                raise ValueError("task [%s] cannot be added: it is in the nullspace of a full joint task" % task.getTaskName())
            if task.task_type == "joint" and task.isFullJointTask():
                cannot_accept_new_tasks = True
        self._joint_limit_avoidance_task = JointLimitAvoidanceTask(robot)
        self._N_constraints = np.eye(robot.dof())
        self._torque_limits = np.full(robot.dof(), np.finfo(np.float64).max)
        for lim in robot.jointLimits():                           # :61-65
            self._torque_limits[lim.joint_index] = lim.effort

    def enableGravityCompensation(self, f): self._enable_gravity_compensation = bool(f)
    def enableTorqueSaturation(self, f): self._enable_torque_saturation = bool(f)

    def enableJointLimitAvoidance(self, f): self._enable_joint_limit_avoidance = bool(f)

    def updateControllerTaskModels(self):                         # :68-77
        n = self._robot.dof()
        N_prec = np.eye(n)
        self._joint_limit_avoidance_task.updateTaskModel(N_prec)  # always, as in the reference (:71)
        self._N_constraints = self._joint_limit_avoidance_task.getTaskAndPreviousNullspace()
        for task in self._tasks:
            task.updateTaskModel(N_prec)
            N_prec = task.getTaskAndPreviousNullspace()

    def computeControlTorques(self):                              # :79-118
        n = self._robot.dof()
        control_torques = np.zeros(n)
        for task in self._tasks:
            control_torques = control_torques + task.computeTorques(control_torques)
        if self._enable_torque_saturation:
            control_torques = np.clip(control_torques, -self._torque_limits, self._torque_limits)
        if self._enable_joint_limit_avoidance:                    # :96-112
            jla = self._joint_limit_avoidance_task.computeTorques(control_torques)
            control_torques = jla + self._N_constraints.T @ control_torques
            if self._enable_torque_saturation:
                control_torques = np.clip(control_torques, -self._torque_limits, self._torque_limits)
        if self._enable_gravity_compensation:
            control_torques = control_torques + self._robot.jointGravityVector()
        return control_torques

    def reinitializeTasks(self):                                  # :120-125
        self._joint_limit_avoidance_task.reInitializeTask()
        for task in self._tasks:
            task.reInitializeTask()
