"""Host build of the 6 x 6 symmetric eigen-solver the blending path runs on the device (csrc/osc_eig6.h: Householder
tridiagonalisation + implicit QL), against LAPACK (numpy.linalg.eigh) on Gram matrices J J^T of task Jacobians:
well conditioned, graded down to the bottom of the reference's blending band, rank deficient and already diagonal."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    out = tmp_path_factory.mktemp("eig6") / "libeig6_probe.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", str(out), os.path.join(HERE, "cpp", "eig6_host_probe.cpp")])
    lib = C.CDLL(str(out))
    lib.eig6_probe.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.eig_probe_m.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def solve(lib, G):
    G = np.ascontiguousarray(G, dtype=np.float64)
    n = G.shape[0]
    Z = np.zeros((n, 6, 6)); d = np.zeros((n, 6))
    lib.eig6_probe(G.ctypes.data, n, Z.ctypes.data, d.ctypes.data)
    return Z, d


def gram_matrices(rng, count, smallest):
    """J = U diag(s) V^T with s from 1 down to `smallest` (6 x 7 Jacobians), returned as J J^T"""
    G = np.zeros((count, 6, 6)); S = np.zeros((count, 6))
    for k in range(count):
        U, _ = np.linalg.qr(rng.standard_normal((6, 6)))
        V, _ = np.linalg.qr(rng.standard_normal((7, 7)))
        s = np.sort(np.concatenate([[1.0], rng.uniform(0.2, 1.0, 3), smallest * rng.uniform(1.0, 3.0, 2)]))[::-1] * rng.uniform(0.3, 3.0)
        J = (U * s) @ V[:, :6].T
        G[k] = J @ J.T
        G[k] = 0.5 * (G[k] + G[k].T)
        S[k] = s
    return G, S


@pytest.mark.parametrize("smallest", [0.3, 0.05, 1e-3])
def test_eigenpairs_against_lapack(probe, smallest):
    rng = np.random.default_rng(int(1e6 * smallest))
    G, S = gram_matrices(rng, 400, smallest)
    Z, d = solve(probe, G)
    for k in range(G.shape[0]):
        lam = np.linalg.eigvalsh(G[k])[::-1]
        order = np.argsort(-d[k])
        scale = lam[0]
        assert np.abs(d[k][order] - lam).max() <= 4e-15 * scale
        # the small eigenvalues keep the relative accuracy the control law needs (sigma enters as 1 / sigma)
        assert (np.abs(np.sqrt(np.maximum(d[k][order], 0)) - np.sqrt(lam)) / np.sqrt(lam)).max() <= 2e-9
        assert np.abs(Z[k].T @ Z[k] - np.eye(6)).max() <= 1e-14
        assert np.abs(G[k] @ Z[k] - Z[k] * d[k]).max() <= 1e-14 * scale


def test_degenerate_inputs(probe):
    rng = np.random.default_rng(5)
    G = np.zeros((5, 6, 6))
    G[0] = np.diag([3.0, 1.0, 2.0, 0.5, 0.0, 4.0])                 # already diagonal, one zero
    J = rng.standard_normal((6, 3)); G[1] = J @ J.T                 # rank three
    G[2] = np.eye(6) * 2.5                                          # multiple eigenvalue
    v = rng.standard_normal(6); G[3] = np.outer(v, v)               # rank one
    G[4] = np.zeros((6, 6))
    Z, d = solve(probe, G)
    for k in range(5):
        lam = np.linalg.eigvalsh(G[k])
        assert np.allclose(np.sort(d[k]), lam, atol=1e-14 * max(1.0, np.abs(lam).max()))
        assert np.abs(Z[k].T @ Z[k] - np.eye(6)).max() <= 1e-14
        assert np.abs(G[k] @ Z[k] - Z[k] * d[k]).max() <= 1e-14 * max(1.0, np.abs(lam).max())


@pytest.mark.parametrize("M", [4, 7, 8])
def test_other_sizes_projector_like_matrices(probe, M):
    """the rolled general path takes the range basis of the joint task's projected Jacobian from sym_eig<number of joints>: Gram
    matrices of null-space projectors (eigenvalues 0 and >= 1, repeated) and of random matrices"""
    rng = np.random.default_rng(M)
    mats = []
    for k in range(200):
        r = int(rng.integers(1, M))
        J = rng.standard_normal((r, M)); Mi = np.linalg.inv(np.diag(rng.uniform(0.5, 2.0, M)) + 0.1 * np.ones((M, M)))
        Nn = np.eye(M) - Mi @ J.T @ np.linalg.inv(J @ Mi @ J.T) @ J          # dynamically consistent null-space projector
        B = Nn if k % 2 == 0 else rng.standard_normal((M, M))
        mats.append(B @ B.T)
    G = np.ascontiguousarray(np.array([0.5 * (g + g.T) for g in mats]))
    Z = np.zeros((len(mats), M, M)); d = np.zeros((len(mats), M))
    probe.eig_probe_m(M, G.ctypes.data, len(mats), Z.ctypes.data, d.ctypes.data)
    for k in range(len(mats)):
        lam = np.linalg.eigvalsh(G[k])
        scale = max(1.0, lam[-1])
        assert np.abs(np.sort(d[k]) - lam).max() <= 1e-13 * scale
        assert np.abs(Z[k].T @ Z[k] - np.eye(M)).max() <= 1e-13
        assert np.abs(G[k] @ Z[k] - Z[k] * d[k]).max() <= 1e-13 * scale
