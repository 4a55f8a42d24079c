"""Kinematics/dynamics stage of the CUDA path vs the sai-model restatement (north_star (a))."""
import numpy as np
import pytest

from oracle.robots import make_chain
from oracle.sai_model import SaiModel
from tests.osc_testlib import TASK_POINTS, rel_err, sample_states

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("robot_name", ["panda", "rrrr", "puma_like", "panda_sliding_base"])
def test_model_stage_matches_oracle(robot_name):
    import sai_primitives_b200 as sp
    N = 64
    q, dq, _ = sample_states(robot_name, N)
    robot = sp.BatchedRobot(robot_name, N)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    link, pt = TASK_POINTS[robot_name]
    out = robot.evalModel(link, pt)
    model = SaiModel(make_chain(robot_name))
    for i in range(N):
        model.setQ(q[i]); model.setDq(dq[i]); model.updateModel()
        assert np.abs(out["M"][i] - model.M()).max() <= 1e-12 * np.abs(model.M()).max()
        assert np.abs(out["J"][i] - model.JWorldFrame(link, pt)).max() <= 1e-13
        assert np.abs(out["x"][i] - model.positionInWorld(link, pt)).max() <= 1e-13
        assert np.abs(out["R"][i] - model.rotationInWorld(link)).max() <= 1e-13
        g = model.jointGravityVector()
        assert np.abs(out["g"][i] - g).max() <= 1e-12 * max(np.abs(g).max(), 1.0)


def test_model_stage_with_base_transform():
    import sai_primitives_b200 as sp
    N = 16
    q, dq, _ = sample_states("panda", N)
    th = 0.3
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]]) @ \
        np.array([[1, 0, 0], [0, np.cos(0.2), -np.sin(0.2)], [0, np.sin(0.2), np.cos(0.2)]])
    t = np.array([0.1, -0.2, 0.3])
    robot = sp.BatchedRobot("panda", N, T_world_robot=(R, t))
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    out = robot.evalModel("end-effector", (0, 0, 0.07))
    model = SaiModel(make_chain("panda"), T_world_robot=(R, t))
    for i in range(N):
        model.setQ(q[i]); model.updateModel()
        assert np.abs(out["J"][i] - model.JWorldFrame("end-effector", (0, 0, 0.07))).max() <= 1e-13
        assert np.abs(out["x"][i] - model.positionInWorld("end-effector", (0, 0, 0.07))).max() <= 1e-13
        assert np.abs(out["g"][i] - model.jointGravityVector()).max() <= 1e-11
        assert np.abs(out["M"][i] - model.M()).max() <= 1e-12 * np.abs(model.M()).max()
