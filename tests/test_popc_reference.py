"""SURVEY.md row a12 pinned to the reference's own code.  tests/golden/popc_reference.npz holds the outputs of
/root/reference/src/helper_modules/POPCExplicitForceControl.cpp (compiled where it lies against oracle/eigen_standin,
oracle/Makefile; generator: tests/golden/generate_popc_reference.py) on a seeded 1,500-step sequence with a passive phase,
an active phase (Rc down to 0.84), a disable/enable cycle and a re-initialisation."""
import ctypes as C
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "popc_reference.npz"))
TOL = 1e-12   # sums of three products: the numpy and CUDA evaluations may round differently in the last bit


def _events(k, obj_disable_enable, obj_reinit):
    if GOLD["events"][k] == 1:
        obj_disable_enable()
    elif GOLD["events"][k] == 2:
        obj_reinit()


def test_numpy_restatement_reproduces_the_reference_code():
    from oracle import primitives as OP
    for key, kv in (("out", GOLD["kv"]), ("out_iso", 10.0 * np.eye(3))):
        c = OP.POPCExplicitForceControl(float(GOLD["dt"]))
        worst = 0.0
        for k in range(GOLD["fd"].shape[0]):
            _events(k, lambda: (c.disable(), c.enable()), c.reInitialize)
            o = c.computePassivitySaturatedForce(GOLD["fd"][k], GOLD["fs"][k], GOLD["vcl"][k], GOLD["vr"][k], kv, float(GOLD["kff"]))
            worst = max(worst, np.abs(o - GOLD[key][k]).max() / max(1.0, np.abs(GOLD[key][k]).max()))
        assert worst < TOL, (key, worst)


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libpopc_ref.so")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_fixture_is_what_the_reference_code_produces_here():
    from tests.golden.generate_popc_reference import make_inputs, run_reference
    inp = make_inputs()
    assert np.array_equal(run_reference(inp), GOLD["out"])
    inp["kv"] = 10.0 * np.eye(3)
    assert np.array_equal(run_reference(inp), GOLD["out_iso"])


@pytest.mark.gpu
def test_cuda_popc_step_reproduces_the_reference_code():
    """the device code of the POPC step (csrc/osc_tasks.cuh: popc_step), driven through osc_debug_popc_sequence"""
    import sai_primitives_b200 as sp
    from sai_primitives_b200 import capi
    N = 3
    robot = sp.BatchedRobot("panda", N)
    robot.setQ(np.zeros((N, 7))); robot.setDq(np.zeros((N, 7))); robot.updateModel()
    mft = sp.MotionForceTask(robot, "end-effector", (np.eye(3), np.array([0, 0, 0.07])))
    sp.RobotController(robot, [mft])
    mft.enablePassivity()
    lib = capi.load_library()
    K = GOLD["fd"].shape[0]
    ev = [k for k in range(K) if GOLD["events"][k] != 0] + [K]
    out = np.zeros((K, 3))
    start = 0
    for stop in ev:                     # segments between events; every event of the sequence re-initialises the POPC state
        if stop > start:
            seg = [np.ascontiguousarray(GOLD[n][start:stop]) for n in ("fd", "fs", "vcl", "vr")]
            o = np.zeros((stop - start, 3))
            rc = lib.osc_debug_popc_sequence(robot.handle, mft.task_id, stop - start, *[capi.host_ptr(a) for a in seg], 10.0, float(GOLD["kff"]), capi.host_ptr(o))
            assert rc == 0, lib.osc_last_error(robot.handle)
            out[start:stop] = o
        if stop < K:
            mft.disablePassivity(); mft.enablePassivity()        # disable() and reInitialize() both reset the POPC state
        start = stop
    err = np.abs(out - GOLD["out_iso"]).max(axis=1) / np.maximum(1.0, np.abs(GOLD["out_iso"]).max(axis=1))
    assert err.max() < TOL, (int(err.argmax()), err.max())
    assert (robot.status() & capi.STATUS_POPC_OVERFLOW).sum() == 0
