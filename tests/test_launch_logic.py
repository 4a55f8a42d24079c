"""Host-side decisions of the kernel launcher (csrc/osc_launch.h), compiled for the host: which hierarchies the specialised
fused kernel -- and with it the optional FP32 mode -- covers (`cycle_spec_eligible`), and when a control cycle takes the split
blending path (`blend_split_selected`).  No GPU needed: the functions only look at the program struct."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    out = tmp_path_factory.mktemp("launch") / "liblaunch_probe.so"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", CUDA_INC, "-o", str(out), os.path.join(HERE, "cpp", "launch_logic_probe.cpp")])
    lib = C.CDLL(str(out))
    lib.min_eig.restype = C.c_double
    lib.min_eig.argtypes = [C.c_void_p]
    return lib


@pytest.mark.parametrize("case,eligible,motion", [
    (0, True, True),      # Panda-like chain, full task under pure motion control + full joint task
    (1, True, False),     # force space: specialised kinematics, general control law
    (2, False, False),    # joint axis not z
    (3, False, False),    # prismatic joint
    (4, False, False),    # joint-task velocity saturation
    (5, False, False),    # rank >= 2 bounded-inertia update possible
    (6, True, False),     # motion-force-task velocity saturation
    (7, False, False),    # batch too large for 32-bit element indices
    (8, True, False),     # closed-loop force control
])
def test_specialised_kernel_gate(probe, case, eligible, motion):
    r = probe.spec_case(case)
    assert bool(r & 1) == eligible and bool(r & 2) == motion


@pytest.mark.parametrize("case,selected", [(0, True), (1, False), (2, False), (3, False), (4, False), (5, False), (6, False), (7, False), (8, False)])
def test_split_blending_path_gate(probe, case, selected):
    assert bool(probe.split_case(case)) == selected


def test_smallest_principal_moment(probe):
    rng = np.random.default_rng(3)
    for _ in range(200):
        A = rng.standard_normal((3, 3)); S = A @ A.T + 0.01 * np.eye(3)
        I6 = np.array([S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2]])
        assert abs(probe.min_eig(I6.ctypes.data) - np.linalg.eigvalsh(S)[0]) < 1e-10 * max(1.0, np.linalg.eigvalsh(S)[-1])
    d = np.array([0.3, 0.0, 0.0, 0.2, 0.0, 0.5])
    assert abs(probe.min_eig(d.ctypes.data) - 0.2) < 1e-15
