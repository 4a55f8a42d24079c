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


def _active_then_passive_sequence(K=3000, seed=3):
    """a contact that stays ACTIVE (observer + correction <= 0, so the reference's window never shrinks) for 2000 cycles with
    power samples of both signs, then turns passive"""
    rng = np.random.default_rng(seed)
    vcl = rng.normal(0, 0.4, (K, 3)); vr = rng.normal(0, 0.02, (K, 3))
    fd = np.tile([0.0, 0.0, -5.0], (K, 1))
    sign = np.where(np.arange(K) < 2000, -1.0, 1.0)[:, None]
    fs = fd + sign * 1.5 * vcl * (1.0 + 0.9 * np.sin(np.arange(K) / 7.0))[:, None] + rng.normal(0, 0.6, (K, 3))
    return fd, fs, vcl, vr


def _reference_popc(fd, fs, vcl, vr, kv, kff, dt=0.001):
    from oracle import primitives as OP
    c = OP.POPCExplicitForceControl(dt); c.enable()
    out = np.zeros_like(fd); longest = 0
    for k in range(fd.shape[0]):
        out[k] = c.computePassivitySaturatedForce(fd[k], fs[k], vcl[k], vr[k], kv * np.eye(3), kff)
        longest = max(longest, len(c._PO_buffer_window))
    return out, longest


@pytest.mark.gpu
@pytest.mark.parametrize("capacity", [4096, 512])
def test_cuda_popc_ring_capacity(capacity):
    """POPCExplicitForceControl.cpp:47-61 keeps every power sample while the observer is <= 0 (unbounded std::queue); the device
    keeps `capacity` samples per robot.  A ring at least as long as the longest active episode reproduces the reference
    exactly; a shorter one raises OSC_STATUS_POPC_OVERFLOW, is identical until the ring first overflows, and deviates afterwards
    (positive samples older than the ring are subtracted from the observer early) -- by how much is printed and bounded here."""
    import sai_primitives_b200 as sp
    from sai_primitives_b200 import capi
    fd, fs, vcl, vr = _active_then_passive_sequence()
    ref, longest = _reference_popc(fd, fs, vcl, vr, 10.0, 0.95)
    assert longest > 2000                                  # the reference's queue really grew past any fixed window
    robot = sp.BatchedRobot("panda", 2)
    robot.setQ(np.zeros((2, 7))); robot.setDq(np.zeros((2, 7))); robot.updateModel()
    mft = sp.MotionForceTask(robot, "end-effector", (np.eye(3), np.array([0, 0, 0.07])))
    sp.RobotController(robot, [mft])
    mft.enablePassivity(ring_capacity=capacity)
    lib = capi.load_library()
    out = np.zeros_like(fd)
    rc = lib.osc_debug_popc_sequence(robot.handle, mft.task_id, fd.shape[0], *[capi.host_ptr(np.ascontiguousarray(a)) for a in (fd, fs, vcl, vr)],
                                     10.0, 0.95, capi.host_ptr(out))
    assert rc == 0, lib.osc_last_error(robot.handle)
    overflow = (robot.status() & capi.STATUS_POPC_OVERFLOW) != 0
    # out = Rc vcl - kv vr  ->  the Rc the device used, per cycle
    rc_dev = ((out + 10.0 * vr) * vcl).sum(axis=1) / (vcl * vcl).sum(axis=1)
    rc_ref = ((ref + 10.0 * vr) * vcl).sum(axis=1) / (vcl * vcl).sum(axis=1)
    if capacity >= longest:
        assert not overflow.any()
        assert np.abs(out - ref).max() < TOL * max(1.0, np.abs(ref).max())
    else:
        assert overflow.all()
        assert np.abs(rc_dev - rc_ref)[:capacity].max() < 1e-9    # identical until the ring first overflows
        assert 0.05 < np.abs(rc_dev - rc_ref).max() <= 1.0         # and really different afterwards: parity is lost, as the bit says
        print("POPC ring %d vs unbounded queue: max |Rc - Rc_ref| = %.3f after overflow" % (capacity, np.abs(rc_dev - rc_ref).max()))
