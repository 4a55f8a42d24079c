"""Golden vectors from the REFERENCE'S OWN CODE for SURVEY.md row a12 (POPC passivity observer / controller).

oracle/_ref/libpopc_ref.so is /root/reference/src/helper_modules/POPCExplicitForceControl.cpp itself, compiled where it lies
against oracle/eigen_standin (oracle/Makefile).  This script drives it with a seeded 1,500-step input sequence that goes
through a passive phase (window pops), an active phase (Rc drops below 1), a disable/enable cycle and a relaxation phase,
and writes inputs and outputs to tests/golden/popc_reference.npz.  Run in the build container (needs /root/reference):
    make -C oracle && python tests/golden/generate_popc_reference.py
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def load_ref():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpopc_ref.so"))
    lib.popc_ref_create.restype = C.c_void_p
    lib.popc_ref_create.argtypes = [C.c_double]
    lib.popc_ref_destroy.argtypes = [C.c_void_p]
    lib.popc_ref_enable.argtypes = [C.c_void_p, C.c_int]
    lib.popc_ref_reinitialize.argtypes = [C.c_void_p]
    P = C.POINTER(C.c_double)
    lib.popc_ref_step.argtypes = [C.c_void_p, P, P, P, P, P, C.c_double, P]
    return lib


def make_inputs(K=1500, seed=20261018):
    g = np.random.default_rng(seed)
    fd = np.tile([0.0, 0.0, -5.0], (K, 1))
    fs = fd + g.normal(0, 1.0, (K, 3))
    vcl = g.normal(0, 2.0, (K, 3))
    vr = g.normal(0, 0.02, (K, 3))
    act = slice(420, 900)                                   # active phase: the velocity follows the commanded force
    vr[act] = 0.05 * (0.95 * fd[act] + vcl[act]) + g.normal(0, 0.01, (act.stop - act.start, 3))
    kv = np.diag([10.0, 10.0, 10.0])
    kv[0, 1] = 0.5                                          # not symmetric on purpose: pins the row-major convention
    events = np.zeros(K, dtype=np.int32)                    # 1: disable then enable before the step, 2: reInitialize
    events[0] = 1
    events[1100] = 1
    events[1300] = 2
    return dict(fd=fd, fs=fs, vcl=vcl, vr=vr, kv=kv, kff=np.float64(0.95), dt=np.float64(0.001), events=events)


def run_reference(inp):
    lib = load_ref()
    h = lib.popc_ref_create(float(inp["dt"]))
    P = C.POINTER(C.c_double)
    K = inp["fd"].shape[0]
    out = np.zeros((K, 3))
    kv = np.ascontiguousarray(inp["kv"])
    for k in range(K):
        if inp["events"][k] == 1:
            lib.popc_ref_enable(h, 0); lib.popc_ref_enable(h, 1)
        elif inp["events"][k] == 2:
            lib.popc_ref_reinitialize(h)
        a = [np.ascontiguousarray(inp[n][k]) for n in ("fd", "fs", "vcl", "vr")]
        o = np.zeros(3)
        lib.popc_ref_step(h, *[x.ctypes.data_as(P) for x in a], kv.ctypes.data_as(P), float(inp["kff"]), o.ctypes.data_as(P))
        out[k] = o
    lib.popc_ref_destroy(h)
    return out


if __name__ == "__main__":
    inp = make_inputs()
    out = run_reference(inp)
    rc = (out + inp["vr"] @ inp["kv"].T) / np.where(np.abs(inp["vcl"]) > 1e-3, inp["vcl"], np.nan)   # Rc per step, for the summary only
    rc = np.nanmedian(rc, axis=1)
    print("steps %d, Rc min %.4f, steps with Rc < 0.999: %d" % (out.shape[0], np.nanmin(rc), int((rc < 0.999).sum())))
    iso = dict(inp)                                         # the form MotionForceTask uses: kv_force * Identity (MotionForceTask.h:308)
    iso["kv"] = 10.0 * np.eye(3)
    out_iso = run_reference(iso)
    np.savez_compressed(os.path.join(HERE, "popc_reference.npz"), out=out, out_iso=out_iso, **inp)
