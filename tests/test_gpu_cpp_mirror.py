"""The C++ host mirror (include/sai_b200/sai_primitives_batched.hpp) drives the same C ABI: build the demo
with g++, run it on the GPU, compare with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from tests.osc_testlib import REL_TOL, TASK_POINTS, OracleBatch, rel_err, sample_states

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_cpp_mirror_example05(tmp_path):
    exe = tmp_path / "host_mirror_demo"
    libdir = os.path.join(ROOT, "sai_primitives_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_demo.cpp"), "-o", str(exe),
                           "-L", libdir, "-lsai_b200_osc", "-Wl,-rpath," + libdir])
    N = 64
    q, dq, _ = sample_states("panda", N, min_sigma_ratio=0.08)
    state = tmp_path / "state.bin"
    with open(state, "wb") as f:
        f.write(np.ascontiguousarray(q.T).tobytes()); f.write(np.ascontiguousarray(dq.T).tobytes())
    out = tmp_path / "tau.bin"
    r = subprocess.run([str(exe), str(state), str(N), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "0 unhandled, 0 on the singular path" in r.stdout
    assert float(r.stdout.strip().split()[-1]) < 1e-9       # a full joint task closes the hierarchy: nothing is left of the null space
    tau = np.fromfile(out, dtype=np.float64).reshape(7, N).T
    link, pt = TASK_POINTS["panda"]
    ob = OracleBatch("panda", N); ob.set_state(q, dq)
    omft = ob.add_mft(link, (np.eye(3), np.array(pt))); ojt = ob.add_jt(); ob.finalize()
    for a, b in zip(omft, ojt):
        a.setGoalLinearVelocity((0.01, -0.02, 0.03)); b.setGoalVelocity(np.full(7, 0.1))
    assert rel_err(tau, ob.cycle()).max() < REL_TOL
