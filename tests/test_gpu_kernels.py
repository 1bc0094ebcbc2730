"""Stage-level parity of the CUDA kernels against the oracle (through the C ABI)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _t(x, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _gemm(mode, A, B, B2, m_begin, m_count, n_begin, n_count, kexp, dev):
    import torch
    from cellregmap_b200 import _lib
    At, Bt = _t(A, dev), _t(B, dev)
    B2t = _t(B2, dev) if B2 is not None else None
    out = torch.full((n_count, m_count + 2), -7.0, dtype=torch.float64, device=dev)
    _lib.call("crm_gemm", mode, _p(At), A.shape[1], A.shape[1], _p(Bt), B.shape[1], B.shape[1],
              _p(B2t) if B2t is not None else ctypes.c_void_p(0), 0 if B2 is None else B2.shape[1], 0 if B2 is None else B2.shape[1],
              A.shape[0], m_begin, m_count, n_begin, n_count, _p(out), m_count + 2, kexp, ctypes.c_void_p(0))
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert np.all(o[:, m_count:] == -7.0)      # nothing written outside the requested block
    return o[:, :m_count]


def _int8_split(X, G, route, dev):
    import torch
    from cellregmap_b200 import _lib
    Xt, Gt = _t(X, dev), _t(G, dev)
    cols, B = X.shape[1], G.shape[1]
    out = torch.full((B, cols + 3), -7.0, dtype=torch.float64, device=dev)
    flags = (ctypes.c_int32 * 2)()
    ms = ctypes.c_float(0.0)
    _lib.call("crm_int8_split_gemm", _p(Xt), cols, cols, _p(Gt), B, B, X.shape[0], route, _p(out), cols + 3, flags, ctypes.byref(ms), ctypes.c_void_p(0))
    o = out.cpu().numpy()
    assert np.all(o[:, cols:] == -7.0)
    return o[:, :cols], (flags[0], flags[1]), ms.value


def _eigh_batched(mats, dev):
    import torch
    from cellregmap_b200 import _lib
    batch, n, _ = mats.shape
    At = _t(mats, dev)
    W = torch.empty((batch, n), dtype=torch.float64, device=dev)
    V = torch.empty((batch, n, n), dtype=torch.float64, device=dev)
    q = (ctypes.c_double * batch)()
    ms = ctypes.c_float(0.0)
    _lib.call("crm_eigh_batched", _p(At), n, batch, _p(W), _p(V), q, ctypes.byref(ms), ctypes.c_void_p(0))
    return W.cpu().numpy(), V.cpu().numpy(), np.array(list(q)), ms.value


def _eig_test_matrices(n, rng):
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    X = rng.standard_normal((3 * n, n))
    mats = [X.T @ X / n,                                                                   # generic positive definite
            3.0 * np.eye(n),                                                               # fully degenerate
            (q * np.repeat(rng.uniform(0.5, 2.0, (n + 7) // 8), 8)[:n]) @ q.T,              # clusters of 8 equal eigenvalues
            (q[:, : n // 3] * rng.uniform(1.0, 5.0, n // 3)) @ q[:, : n // 3].T,            # rank n/3, null space of dimension 2n/3
            (q * np.concatenate([np.full(n // 2, 1.0), 1.0 + 1e-9 * np.arange(n - n // 2)])) @ q.T,   # very close, not equal
            np.diag(rng.uniform(-1.0, 1.0, n))]                                            # already diagonal (every reflector trivial)
    return np.stack([(m + m.T) / 2 for m in mats])


@pytest.mark.parametrize("n", [2, 5, 64, 257, 1020])
def test_eigh_batched(cuda_device, n):
    """Set-up eigensolver (eig.cuh): eigenvalues against LAPACK, orthonormality and residuals of the eigenvectors, including
    degenerate, clustered and rank-deficient spectra."""
    rng = np.random.default_rng(n)
    mats = _eig_test_matrices(n, rng) if n >= 24 else np.stack([(lambda a: (a + a.T) / 2)(rng.standard_normal((n, n))) for _ in range(4)])
    W, V, q, ms = _eigh_batched(mats, cuda_device)
    for b in range(mats.shape[0]):
        A = mats[b]
        scale = max(np.abs(np.linalg.eigvalsh(A)).max(), 1e-300)
        Vb = V[b].T                                   # V[b][t] = eigenvector t -> columns
        assert np.isfinite(q[b]) and q[b] < 1e-11, (b, q[b])
        assert np.all(np.diff(W[b]) >= -1e-14 * scale)
        assert np.max(np.abs(W[b] - np.linalg.eigvalsh(A))) < 1e-13 * n * scale, b
        assert np.max(np.abs(Vb.T @ Vb - np.eye(n))) < 1e-12, b
        assert np.max(np.abs(A @ Vb - Vb * W[b])) < 1e-12 * n * scale, b


@pytest.mark.parametrize("n,cols,B", [(16, 8, 5), (1000, 130, 37), (4099, 300, 260), (20000, 129, 515), (700, 1200, 3)])
def test_int8_split_contraction(cuda_device, n, cols, B):
    """K0: the fused tcgen05 kernel equals the cuBLASLt + recombination route bit for bit, and both equal the exact
    contraction (integer-valued X: every product and sum is exactly representable)."""
    rng = np.random.default_rng(n + cols)
    G = rng.integers(0, 3, size=(n, B)).astype(np.float64)
    X = rng.standard_normal((n, cols)) * np.exp(rng.uniform(-8, 8, size=cols))
    X[:, min(3, cols - 1)] = 0.0                              # an all-zero column (empty exponent)
    c_mma, f0, _ = _int8_split(X, G, 0, cuda_device)
    c_lt, f1, _ = _int8_split(X, G, 1, cuda_device)
    c_pair, f2, _ = _int8_split(X, G, 2, cuda_device)
    assert f0 == f1 == f2 == (0, 2)
    assert np.array_equal(c_mma, c_lt)
    assert np.array_equal(c_pair, c_lt)
    ref = G.T @ X
    scale = np.abs(G).T @ np.abs(X) + 1e-300
    assert np.max(np.abs(c_mma - ref) / scale) < 1e-14
    Xi = np.round(rng.standard_normal((n, cols)) * 1000.0)   # integers: exact result
    c_i, _, _ = _int8_split(Xi, G, 0, cuda_device)
    assert np.array_equal(c_i, G.T @ Xi)


@pytest.mark.parametrize("K,M,N", [(1, 2, 2), (37, 130, 6), (1000, 222, 300), (4099, 64, 129)])
def test_gemm_plain(cuda_device, K, M, N):
    rng = np.random.default_rng(K + M + N)
    Mp, Np = M + (M & 1), N + (N & 1)
    A = rng.standard_normal((K, Mp)); B = rng.standard_normal((K, Np))
    got = _gemm(0, A, B, None, 0, M, 0, N, 1, cuda_device)
    want = B[:, :N].T @ A[:, :M]
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12 * np.sqrt(K))


def test_gemm_sub_blocks(cuda_device):
    rng = np.random.default_rng(5)
    A = rng.standard_normal((513, 300)); B = rng.standard_normal((513, 400))
    got = _gemm(0, A, B, None, 37, 150, 11, 260, 1, cuda_device)
    want = B[:, 11:271].T @ A[:, 37:187]
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-11)


def test_gemm_product(cuda_device):
    rng = np.random.default_rng(6)
    A = rng.standard_normal((777, 232)); G = rng.integers(0, 3, (777, 150)).astype(float)
    got = _gemm(1, A, G, G, 0, 231, 0, 149, 1, cuda_device)
    want = (G[:, :149] ** 2).T @ A[:, :231]
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-11)


@pytest.mark.parametrize("k", [1, 3, 10, 20, 41])
def test_gemm_expand(cuda_device, k):
    rng = np.random.default_rng(k)
    K, M, S = 1531, 200, 57
    kexp = 1 + k
    pitch = (kexp + 2) & ~1
    while pitch % 16 not in (4, 12):
        pitch += 2
    A = rng.standard_normal((K, M)); G = rng.integers(0, 3, (K, S + (S & 1))).astype(float)
    E = rng.standard_normal((K, k))
    Eext = np.zeros((K, pitch)); Eext[:, 0] = 1.0; Eext[:, 1:kexp] = E
    got = _gemm(2, A, G, Eext, 0, M, 0, S * kexp, kexp, cuda_device)
    cols = []
    for s in range(S):
        cols.append(G[:, s]); cols += [G[:, s] * E[:, j] for j in range(k)]
    want = np.stack(cols, axis=1).T @ A
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-11)


def _davies_gpu(Q, lams, dev):
    import torch
    from cellregmap_b200 import _lib
    count = len(Q)
    ld = max(len(l) for l in lams)
    lam = np.zeros((count, ld)); nlam = np.zeros(count, np.int32)
    for i, l in enumerate(lams):
        lam[i, :len(l)] = l; nlam[i] = len(l)
    Qt, lt, nt = _t(np.asarray(Q, float), dev), _t(lam, dev), _t(nlam, dev)
    pv = torch.empty(count, dtype=torch.float64, device=dev); liu = torch.empty_like(pv)
    ifault = torch.empty(count, dtype=torch.int32, device=dev); conv = torch.empty_like(ifault)
    trace = torch.empty((count, 8), dtype=torch.float64, device=dev)
    _lib.call("crm_davies_pvalues", _p(Qt), _p(lt), _p(nt), ld, count, 10000, 1e-6, _p(pv), _p(liu), _p(ifault), _p(conv), _p(trace), ctypes.c_void_p(0))
    torch.cuda.synchronize()
    return pv.cpu().numpy(), liu.cpu().numpy(), ifault.cpu().numpy(), conv.cpu().numpy(), trace.cpu().numpy()


def test_davies_matches_oracle(cuda_device):
    from oracle import chiscore_port as cp
    rng = np.random.default_rng(11)
    Q, lams = [], []
    for i in range(300):
        r = int(rng.integers(1, 25))
        lam = np.sort(rng.gamma(0.7, 1.0, r))[::-1] + 1e-6
        q = lam.sum() * float(rng.choice([0.05, 0.5, 1.0, 2.0, 4.0, 8.0, 16.0, 40.0]))
        Q.append(q); lams.append(lam)
    pv, liu, ifault, conv, trace = _davies_gpu(Q, lams, cuda_device)
    for i in range(len(Q)):
        p, info = cp.pvalue_from_lambda(lams[i], Q[i])
        assert ifault[i] == info["ifault"], (i, ifault[i], info["ifault"])
        assert conv[i] == info["Is_Converged"]
        assert trace[i, 7] == info["trace"][6]            # same number of bound evaluations
        assert trace[i, 2] == info["trace"][1]            # same number of integration terms
        assert abs(np.log10(liu[i]) - np.log10(info["liu_pval"])) < 1e-6
        if p >= 1e-12:
            assert abs(np.log10(pv[i]) - np.log10(p)) <= 1e-4, (i, pv[i], p)
        else:
            assert abs(np.log10(pv[i]) - np.log10(p)) <= 1e-3


def test_davies_edge_cases(cuda_device):
    pv, liu, ifault, conv, _ = _davies_gpu([1.0, 1.0], [np.array([0.5]), np.array([])], cuda_device)
    from scipy.stats import chi2
    assert abs(pv[0] - chi2.sf(2.0, 1)) < 1e-8       # single eigenvalue -> Liu, exact for one chi2_1
    assert np.isnan(pv[1]) and ifault[1] == -1       # no eigenvalue -> flagged


def test_lrt_pvalues(cuda_device):
    from cellregmap_b200 import lrt_pvalues
    from oracle.crm_port import lrt_pvalues as ref
    alt = np.array([-100.0, -99.0, -90.0, -50.0, -100.5, -100.0 + 1e-9, 700.0])
    got = lrt_pvalues(-100.0, alt)
    want = ref(-100.0, alt)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
    for dof in (2, 3, 7, 0.5):         # any number of degrees of freedom (reference :443-469 takes `dof`)
        np.testing.assert_allclose(lrt_pvalues(-100.0, alt, dof=dof), ref(-100.0, alt, dof=dof), rtol=1e-11, atol=0)


@pytest.mark.parametrize("restricted", [True, False])
@pytest.mark.parametrize("c", [1, 3])
def test_lmm_fit_rotated_matches_oracle(cuda_device, restricted, c):
    import torch
    from cellregmap_b200 import _lib
    from oracle.lmm_port import LMM
    from oracle.sugar_port import economic_qs_linear
    rng = np.random.default_rng(100 + c)
    n, r, p, R = 300, 40, 12, 3
    y = rng.standard_normal(n)
    W = np.column_stack([np.ones(n)] + [rng.standard_normal(n) for _ in range(c - 1)])
    G = rng.integers(0, 3, (n, p)).astype(float)
    mp_ = r + (r & 1)
    S = np.zeros((R, mp_)); yr = np.zeros((R, mp_)); Wr = np.zeros((R, c, mp_)); gr = np.zeros((p, R, mp_))
    QSs = []
    for k in range(R):
        Gh = rng.standard_normal((n, r)) * (0.3 + 0.4 * k)
        if k == 0:
            y = y + 0.5 * Gh @ rng.standard_normal(r) / np.sqrt(r)
        QSs.append(economic_qs_linear(Gh, return_q1=False))
    for k in range(R):
        Q0, S0 = QSs[k][0][0], QSs[k][1]
        S[k, :r] = S0; yr[k, :r] = Q0.T @ y; Wr[k, :, :r] = (Q0.T @ W).T; gr[:, k, :r] = (Q0.T @ G).T
    stats = np.concatenate([[y @ y], W.T @ y, (W.T @ W).ravel()])
    dev = cuda_device
    args = [_t(a, dev) for a in (S, yr, Wr, gr, G.T @ y, G.T @ W, (G * G).sum(0), stats)]
    P = c + 1
    lml = torch.empty((p, R), dtype=torch.float64, device=dev); delta = torch.empty_like(lml); scale = torch.empty_like(lml)
    beta = torch.empty((p, R, P), dtype=torch.float64, device=dev)
    nfev = torch.empty((p, R), dtype=torch.int32, device=dev); flags = torch.empty_like(nfev)
    _lib.call("crm_lmm_fit_rotated", *[_p(a) for a in args], r, mp_, R, c, p, float(n), int(restricted),
              _p(lml), _p(delta), _p(scale), _p(beta), _p(nfev), _p(flags), ctypes.c_void_p(0))
    torch.cuda.synchronize()
    lml, delta, scale, beta, nfev = [a.cpu().numpy() for a in (lml, delta, scale, beta, nfev)]
    for s in range(p):
        for k in range(R):
            ref = LMM(y, np.column_stack([W, G[:, s]]), QSs[k], restricted=restricted)
            ref.fit(verbose=False)
            assert abs(lml[s, k] - ref.lml()) <= 1e-6 * abs(ref.lml())
            np.testing.assert_allclose(delta[s, k], ref.delta, rtol=1e-6)
            np.testing.assert_allclose(scale[s, k], ref.scale, rtol=1e-6)
            # beta is a ratio of O(1e-6)-perturbed sums for the weakly determined coefficients of the random design here
            np.testing.assert_allclose(beta[s, k], ref.beta, rtol=1e-5, atol=1e-8)
            # same counting rule as the port (every objective evaluation of bracket + Brent); a last-bit difference in an lml can
            # end Brent one step earlier (the path rule of DESIGN.md section 5), never later by more than that on these cases
            assert nfev[s, k] == ref.nfev - 1 or nfev[s, k] == ref.nfev


def test_liu_params_and_qmin_goldens(cuda_device):
    """Known answers of the reference (cellregmap/test/test_math.py:76-91) through the CUDA ops."""
    from cellregmap_b200._math import liu_params_batch, qmin, qmin_batch, score_statistic_liu_params
    from oracle import math_port as mp
    w = np.array([4.55266277e-09, 3.46249449e-01])
    got = score_statistic_liu_params(0.49961017073389324, w)
    np.testing.assert_allclose(got["pv"], 0.22966744652848403, rtol=1e-7)
    np.testing.assert_allclose(got["mu_q"], 0.34624945394475326, rtol=1e-7)
    np.testing.assert_allclose(got["sigma_q"], 0.48967066729451103, rtol=1e-7)
    np.testing.assert_allclose(got["dof_x"], 1.0, rtol=1e-7)
    params = [{"pv": 0.22966742, "mu_q": 0.34945, "sigma_q": 0.48670, "dof_x": 1.5}, {"pv": 0.65, "mu_q": 0.695, "sigma_q": 0.1, "dof_x": 0.7}]
    np.testing.assert_allclose(qmin(params), [0.5506645025120773, 0.7157125486956082], rtol=1e-7)
    # random batch against the oracle
    rng = np.random.default_rng(0)
    ws = [np.sort(rng.gamma(0.8, 1.0, rng.integers(1, 15)))[::-1] + 1e-6 for _ in range(64)]
    qs = np.array([w.sum() * rng.uniform(0.2, 6.0) for w in ws])
    got = liu_params_batch(qs, ws)
    for i in range(64):
        ref = mp.score_statistic_liu_params(qs[i], ws[i])
        np.testing.assert_allclose(got[i, 0], ref["pv"], rtol=1e-7)      # scipy's ncx2.sf vs the Poisson-mixture form
        np.testing.assert_allclose(got[i, 1:], [ref["mu_q"], ref["sigma_q"], ref["dof_x"]], rtol=1e-9)
    P = got.reshape(8, 8, 4)
    out = qmin_batch(P)
    for b in range(8):
        ref = mp.qmin([{"pv": P[b, r, 0], "mu_q": P[b, r, 1], "sigma_q": P[b, r, 2], "dof_x": P[b, r, 3]} for r in range(8)])
        np.testing.assert_allclose(out[b], ref, rtol=1e-8)
