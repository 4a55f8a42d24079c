"""numpy prototype of the algebra the CUDA fast path uses.

Not the oracle and not the product: it exists so that the reformulation
(whitened coordinates M = L L^T, hierarchy by Householder QR, joint task in the
null space through an orthonormal complement) can be checked against the
literal restatement in oracle/primitives.py on the CPU, without a GPU.

Everything here assumes the non-singular branch of
SingularityHandler::updateTaskModel (reference SingularityHandler.cpp:123-141).
"""
import numpy as np

FULL, BIE, IMP = 0, 1, 2


def householder_append(V, beta, c, X):
    """Householder QR of X[c:, :] (X already rotated by the c existing
    reflectors).  Appends r reflectors acting on rows c.., returns R (r x r)."""
    n, r = X.shape
    X = X.copy()
    R = np.zeros((r, r))
    for j in range(r):
        k = c + j
        x = X[k:, j]
        alpha = -np.copysign(np.linalg.norm(x), x[0] if x[0] != 0 else 1.0)
        v = x.copy()
        v[0] -= alpha
        vv = v @ v
        b = 0.0 if vv == 0 else 2.0 / vv
        V[k:, k] = v
        V[:k, k] = 0
        beta[k] = b
        for jj in range(j, r):
            X[k:, jj] -= b * v * (v @ X[k:, jj])
        R[: j + 1, j] = 0  # filled below
    # after the loop the rotated block is upper triangular in rows c..c+r-1
    R = np.triu(X[c:c + r, :])
    return R


def apply_Qt(V, beta, c, x):
    """x <- H_c ... H_1 x"""
    x = x.copy()
    for k in range(c):
        v = V[:, k]
        x -= beta[k] * v * (v @ x)
    return x


def apply_Q(V, beta, c, x):
    """x <- H_1 ... H_c x"""
    x = x.copy()
    for k in range(c - 1, -1, -1):
        v = V[:, k]
        x -= beta[k] * v * (v @ x)
    return x


def sound_nonsingular(Jp, thr):
    """Sufficient test for s_min/s_max >= thr on the rows of Jp (r x n):
    lambda_max(G) <= (tr G^8)^(1/8); then G - thr^2*hi*I must be positive
    definite (Cholesky succeeds)."""
    G = Jp @ Jp.T
    G2 = G @ G
    G4 = G2 @ G2
    hi = np.sum(G4 * G4) ** 0.125
    try:
        np.linalg.cholesky(G - (thr * thr) * hi * np.eye(G.shape[0]))
        return True
    except np.linalg.LinAlgError:
        return False


def fast_cycle(M, tasks, q, dq, thr_bie_default=0.1):
    """tasks: list of dicts
       MFT: {kind:'mft', B:(6 x r), J0:(6 x n), fstar:(6,), F:(6,), dec:.., bie:..}
       JT : {kind:'jt', S:(k x n), acc:(k,), t:(k,), dec:.., bie:..}
       (fstar/F/t/acc are the control-law outputs, computed by the caller)
    returns (tau_total, ok)"""
    n = M.shape[0]
    L = np.linalg.cholesky(M)
    V = np.zeros((n, n)); beta = np.zeros(n); c = 0
    tau = np.zeros(n)
    ok = True
    for tk in tasks:
        Mb = M.copy()
        for i in range(n):
            if Mb[i, i] < tk["bie"]:
                Mb[i, i] = tk["bie"]
        Lb = np.linalg.cholesky(Mb)
        Jt = tk["B"].T @ tk["J0"] if tk["kind"] == "mft" else tk["S"]
        r = Jt.shape[0]
        m = n - c
        if tk["kind"] == "jt" and r > m:
            # rank-deficient joint task: only the full joint task is handled here
            assert r == n and np.allclose(tk["S"], np.eye(n))
            if m == 0:
                continue  # zero range: task contributes nothing (JointTask.cpp:234-239, 302-306)
            Qp = np.zeros((n, m))
            for a in range(m):
                e = np.zeros(n); e[c + a] = 1.0
                Qp[:, a] = apply_Q(V, beta, c, e)
            W = np.linalg.solve(L.T, Qp)
            K = L @ Qp
            G = W.T @ W
            if np.trace(G @ (K.T @ K)) >= 1e6:
                ok = False
            z_acc = np.linalg.solve(G, W.T @ tk["acc"])
            Minv_tp = np.linalg.solve(L.T, np.linalg.solve(L, tau))
            z_dist = np.linalg.solve(G, W.T @ Minv_tp)
            if tk["dec"] == FULL:
                z_t = np.linalg.solve(G, W.T @ tk["t"])
            elif tk["dec"] == BIE:
                Z = np.linalg.solve(Lb, K)
                H = Z.T @ Z
                z_t = np.linalg.solve(H, np.linalg.solve(G, W.T @ tk["t"]))
            else:
                z_t = W.T @ tk["t"]
            tau = tau + K @ (z_acc + z_t - z_dist)
            continue
        X = np.linalg.solve(L, Jt.T)
        for a in range(r):
            X[:, a] = apply_Qt(V, beta, c, X[:, a])
        Xp = X.copy(); Xp[:c, :] = 0
        JpT = np.zeros((n, r))
        for a in range(r):
            JpT[:, a] = L @ apply_Q(V, beta, c, Xp[:, a])
        thr = 0.06 if tk["kind"] == "mft" else 1e-3
        if not sound_nonsingular(JpT.T, thr):
            ok = False
        R = householder_append(V, beta, c, X)
        c += r

        def lam_full(y):
            return np.linalg.solve(R, np.linalg.solve(R.T, y))

        def lam_mod(y):
            if tk["dec"] == FULL:
                return lam_full(y)
            if tk["dec"] == BIE:
                Wb = np.linalg.solve(Lb, JpT)
                return np.linalg.solve(Wb.T @ Wb, y)
            return y

        if tk["kind"] == "mft":
            yf = tk["B"].T @ tk["fstar"]; yF = tk["B"].T @ tk["F"]
            tau = tau + JpT @ (lam_mod(yf) + yF)
        else:
            Minv_tp = np.linalg.solve(L.T, np.linalg.solve(L, tau))
            tau = tau + JpT @ (lam_full(tk["acc"]) + lam_mod(tk["t"]) - lam_full(tk["S"] @ Minv_tp))
    return tau, ok
