"""Shared helpers of the parity tests: seeded synthetic inputs and
oracle-vs-CUDA comparison.  The oracle (oracle/) is used here ONLY as the checker.

Two checkers exist:
  "reference"  oracle/_ref/libsai_ref_orient.so -- the reference's OWN JointTask / MotionForceTask / SingularityHandler /
               JointLimitAvoidanceTask / RobotController / POPC sources compiled where they lie (oracle/sai_ref.py);
  "numpy"      oracle/primitives.py -- the statement-by-statement restatement.
`OracleBatch(...)` hands out the reference whenever the library is present (it is built in this container and travels to the
GPU box) and the numpy restatement otherwise; OSC_ORACLE=numpy|reference forces one."""
import os

import numpy as np

from oracle import primitives as OP
from oracle.robots import make_chain
from oracle.sai_model import SaiModel

SEED = 1234
REL_TOL = 1e-9   # BASELINE.json north_star: torques within 1e-9 relative in FP64

TASK_POINTS = {
    "panda": ("end-effector", (0.0, 0.0, 0.07)),          # examples/05-using_robot_controller/...cpp:111-113
    "panda_sliding_base": ("end-effector", (0.0, 0.0, 0.07)),
    "rrrr": ("link4", (0.5, 0.0, 0.0)),                    # examples/11-planar_robot_controller/...cpp:105-115
    "puma_like": ("end-effector", (0.0, 0.0, 0.0)),
}


def rng_for(robot_index, stream=0):
    """counter-based generator: one Philox stream per robot index"""
    return np.random.Generator(np.random.Philox(key=SEED + stream, counter=[0, 0, 0, int(robot_index)]))


def rot_exp(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def sample_states(robot_name, n_robots, min_sigma_ratio=None, dirs=None, first_index=0, max_tries=400):
    """q ~ U(lo + 0.1 range, hi - 0.1 range), dq ~ U(-1, 1); optionally rejects states whose task
    Jacobian has s_{r-1}/s_0 < min_sigma_ratio (so the non-singular branch is exercised)."""
    ch = make_chain(robot_name)
    model = SaiModel(ch)
    link, pt = TASK_POINTS[robot_name]
    n = ch.n
    q = np.zeros((n_robots, n)); dq = np.zeros((n_robots, n))
    rejected = 0
    for i in range(n_robots):
        g = rng_for(first_index + i)
        for _ in range(max_tries):
            qi = ch.q_lower + (0.1 + 0.8 * g.random(n)) * (ch.q_upper - ch.q_lower)
            dqi = g.uniform(-1.0, 1.0, n)
            if min_sigma_ratio is None:
                break
            model.setQ(qi); model.updateKinematics()
            J = model.J(link, pt)
            if dirs is not None:
                J = dirs.T @ J
            s = np.linalg.svd(J, compute_uv=False)
            r = J.shape[0] if dirs is not None else min(J.shape)
            if s[r - 1] / s[0] >= min_sigma_ratio:
                break
            rejected += 1
        else:
            raise RuntimeError("could not sample a non-singular state")
        q[i], dq[i] = qi, dqi
    return q, dq, rejected


def sensed_ex09(N, k):
    """sensed wrench of cycle k (0-based) for the ex.09 closed-loop force scenario: goal force (0,0,-5) + noise whose level
    alternates every 60 cycles so that the passivity controller engages and relaxes"""
    F = np.zeros((N, 3)); Mo = np.zeros((N, 3))
    for i in range(N):
        g = rng_for(i * 100003 + k, stream=33)
        F[i] = np.array([0, 0, -5.0]) + g.normal(0, 1.0, 3) * (3.0 if (k // 60) % 2 else 1.0)
        Mo[i] = g.normal(0, 0.1, 3)
    return F, Mo


def rel_err(a, b):
    """per-robot max abs error relative to the robot's reference infinity norm"""
    a = np.asarray(a); b = np.asarray(b)
    scale = np.maximum(np.abs(b).reshape(b.shape[0], -1).max(axis=1), 1e-9)
    return np.abs(a - b).reshape(a.shape[0], -1).max(axis=1) / scale


def oracle_kind():
    forced = os.environ.get("OSC_ORACLE")
    if forced:
        return forced
    from oracle import sai_ref
    return "reference" if sai_ref.available(oriented=True) else "numpy"


def OracleBatch(robot_name, n_robots, T_world_robot=None, kind=None):
    """N independent CPU controllers with the same hierarchy: the reference's compiled control law when available"""
    kind = kind or oracle_kind()
    if kind == "reference":
        from oracle.sai_ref import RefBatch
        return RefBatch(robot_name, n_robots, T_world_robot=T_world_robot, oriented=True)
    return NumpyOracleBatch(robot_name, n_robots, T_world_robot=T_world_robot)


class NumpyOracleBatch:
    """N independent oracle robots with the same hierarchy (the CPU restatement looped over the batch)."""

    def __init__(self, robot_name, n_robots, T_world_robot=None):
        self.chain = make_chain(robot_name)
        self.robots = [SaiModel(self.chain, T_world_robot=T_world_robot) for _ in range(n_robots)]
        self.tasks = [[] for _ in range(n_robots)]
        self.controllers = None

    def set_state(self, q, dq):
        for r, qi, dqi in zip(self.robots, q, dq):
            r.setQ(qi); r.setDq(dqi); r.updateModel()

    def add_mft(self, link, compliant=None, dirs_t=None, dirs_r=None, in_compliant=False, dt=0.001, name="motion_force_task"):
        out = []
        for i, r in enumerate(self.robots):
            t = OP.MotionForceTask(r, link, compliant, dirs_t, dirs_r, task_name=name,
                                   is_force_motion_parametrization_in_compliant_frame=in_compliant, loop_timestep=dt)
            self.tasks[i].append(t); out.append(t)
        return out

    def add_jt(self, S=None, dt=0.001, name="joint_task"):
        out = []
        for i, r in enumerate(self.robots):
            t = OP.JointTask(r, S, task_name=name, loop_timestep=dt)
            self.tasks[i].append(t); out.append(t)
        return out

    def finalize(self):
        self.controllers = [OP.RobotController(r, ts) for r, ts in zip(self.robots, self.tasks)]

    def cycle(self, use_prev=True):
        taus = []
        for r, c, ts in zip(self.robots, self.controllers, self.tasks):
            c.updateControllerTaskModels()
            if use_prev:
                taus.append(c.computeControlTorques())
            else:
                tau = np.zeros(r.dof())
                for t in ts:
                    tau = tau + t.computeTorques()
                taus.append(tau)
        return np.array(taus)
