"""Host-side pieces of the Python mirror that run without a GPU (torch CPU tensors)."""
import numpy as np
import pytest
import torch

from cellregmap_b200 import _cellregmap as api
from oracle import crm_port


@pytest.mark.parametrize("case", ["generic", "duplicate_columns", "one_hot"])
def test_scaled_left_vectors_match_numpy_svd(case):
    """U S of the thin SVD through the triangular factor (reference get_L_values, cellregmap/_cellregmap.py:533-545): same column
    space, same number of kept directions, same E E' as numpy's SVD with the sqrt(eps) cut."""
    rng = np.random.default_rng(3)
    if case == "generic":
        E = rng.standard_normal((400, 9))
    elif case == "duplicate_columns":
        E = np.hstack([rng.standard_normal((300, 4)), np.ones((300, 1)), np.ones((300, 1))])
    else:
        E = np.eye(6)[rng.integers(0, 6, 500)]
        E = np.hstack([E, np.ones((500, 1))])          # intercept + one-hot: rank 6 of 7
    us = api._scaled_left_vectors(torch.from_numpy(E))[0].numpy()
    U, S, _ = np.linalg.svd(E, full_matrices=False)
    keep = S >= api.EPS_SMALL
    assert us.shape == (E.shape[0], int(keep.sum()))
    np.testing.assert_allclose(us @ us.T, (U[:, keep] * S[keep]) @ (U[:, keep] * S[keep]).T, rtol=0, atol=1e-11 * S[0] ** 2)
    np.testing.assert_allclose(np.sort(np.linalg.norm(us, axis=0))[::-1], S[keep], rtol=1e-10)


@pytest.mark.parametrize("case", ["generic", "cond 1e3", "cond 1e6"])
def test_context_map_routes_agree(case, monkeypatch):
    """The context map V with U S = E V comes from the k x k Gram for well-conditioned contexts and from the QR factor otherwise
    (sigma_min / sigma_max < 3e-5); both give an orthogonal V, U S = E V exactly, and the same E E'."""
    rng = np.random.default_rng(8)
    E = rng.standard_normal((600, 6))
    if case != "generic":
        E[:, 5] = E[:, 0] + (1e-3 if case == "cond 1e3" else 1e-6) * E[:, 5]
    Et = torch.from_numpy(E)
    us_a, V_a = api._scaled_left_vectors(Et)
    monkeypatch.setenv("CRM_CONTEXT_SVD", "qr")
    us_b, V_b = api._scaled_left_vectors(Et)
    for us, V in ((us_a, V_a), (us_b, V_b)):
        assert V.shape == (6, 6) and V.flags["C_CONTIGUOUS"]
        np.testing.assert_allclose(V.T @ V, np.eye(6), atol=1e-13)
        np.testing.assert_allclose(us.numpy(), E @ V, rtol=0, atol=1e-12 * np.abs(E).max())
    np.testing.assert_allclose((us_a @ us_a.T).numpy(), (us_b @ us_b.T).numpy(), rtol=0, atol=1e-11 * np.abs(E @ E.T).max())
    # which route ran: the two agree column by column (up to sign) only when both are singular vectors of a well-separated spectrum
    S = np.linalg.svd(E, compute_uv=False)
    assert (S[-1] / S[0] < 3e-5) == (case == "cond 1e6")


def test_l_concat_equals_the_oracle_blocks():
    """sum_i L_i L_i' from the concatenated blocks equals the oracle's get_L_values (signs of the blocks are free)."""
    rng = np.random.default_rng(4)
    E = rng.standard_normal((120, 5))
    hK = rng.standard_normal((120, 3))
    L = api._L_concat(torch.from_numpy(hK), torch.from_numpy(E)).numpy()
    Ls = crm_port.get_L_values(hK, E)
    K_ref = sum(Li @ Li.T for Li in Ls)
    np.testing.assert_allclose(L @ L.T, K_ref, rtol=0, atol=1e-11 * np.abs(K_ref).max())
    assert L.shape == (120, sum(Li.shape[1] for Li in Ls))


def test_integer_genotypes_are_recognised_without_a_gpu():
    """The narrow-transfer route is only taken for integer dtypes of at most 4 bytes; everything else keeps the float64 host route
    (checked on the decision alone: no device here)."""
    for dtype, narrow in ((np.int8, True), (np.uint8, True), (np.int16, True), (np.int32, True), (np.int64, False), (np.float32, False), (np.float64, False)):
        arr = np.zeros((3, 2), dtype=dtype)
        assert (arr.dtype.kind in "iub" and arr.dtype.itemsize <= 4) == narrow


def test_get_L_values_remembers_its_factors():
    """get_L_values returns the reference's list of blocks (values as before) and keeps (hK, E, V) with U S = E V beside it, so that a model
    built from the list can declare the structure of the background; copies and slices are plain lists without it."""
    rng = np.random.default_rng(12)
    E = rng.standard_normal((90, 4))
    hK = rng.standard_normal((90, 3))
    Ls = api.get_L_values(hK, E)
    ref = crm_port.get_L_values(hK, E)
    assert isinstance(Ls, list) and len(Ls) == len(ref) == 4
    for a, b in zip(Ls, ref):
        np.testing.assert_allclose(np.abs(a), np.abs(b), rtol=0, atol=1e-13)
    hK_f, E_f, V = Ls.factors
    assert hK_f.shape == hK.shape and E_f.shape == E.shape and V.shape == (4, 4)
    np.testing.assert_allclose(np.concatenate(Ls, 1), ((E @ V)[:, :, None] * hK[:, None, :]).reshape(90, 12), rtol=0, atol=1e-13)
    assert not hasattr(Ls[:2], "factors") and not hasattr(list(Ls), "factors")
    # wide context matrices (fewer cells than contexts) have no such map
    assert api.get_L_values(rng.standard_normal((3, 2)), rng.standard_normal((3, 5))).factors is None
