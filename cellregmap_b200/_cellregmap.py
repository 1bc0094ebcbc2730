"""Host-side mirror of the reference API (cellregmap/_cellregmap.py) on top of libcrm_b200.

Same names, argument meaning, return shapes and error behaviour as the reference:
`CellRegMap(y, E, W=None, Ls=None, E1=None, hK=None)` (reference :63), `.scan_interaction` (:317),
`.scan_association` (:246), `.scan_association_fast` (:284), `run_interaction` (:547),
`run_association` (:471), `run_association_fast` (:502), `get_L_values` (:533), `lrt_pvalues` (:443),
`compute_maf` (:589).  Inputs may be numpy arrays or torch tensors (CPU or CUDA); results are numpy
arrays like the reference's.  All numerics run in the CUDA library; torch only owns device memory
and streams.
"""
import ctypes
import os
from typing import Optional

import numpy as np
import torch

from . import _lib

EPS_SMALL = float(np.sqrt(np.finfo(float).eps))

# bench.py switches this on to time the rotation kernel with CUDA events inside the library
PROFILE = {"on": False, "rot_ms": 0.0, "rot_flops": 0.0, "rot_launches": 0, "int8_ms": 0.0, "int8_ops": 0.0, "int8_launches": 0}


def _device(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("cellregmap_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _to_dev(x, device, two_d=False):
    """asarray(x, float) onto the device as a C-contiguous float64 tensor."""
    if isinstance(x, torch.Tensor):
        t = x.to(device=device, dtype=torch.float64)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(device)
    if two_d and t.ndim == 1:
        t = t.reshape(-1, 1)
    return t.contiguous()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _scaled_left_vectors(E):
    """(U S, V) of the thin SVD of the tall matrix E with singular values >= sqrt(eps) (numpy_sugar.linalg.economic_svd), up to the
    sign of each column; V (k x r, numpy) is the map U S = E V, None when E is wider than tall.  Through the triangular factor:
    E = Q R, R = Ur S V' (k x k, on the host: one 3 KB read-back).
    torch.linalg.svd on the device runs an iterative cuSOLVER routine with dozens of small launches and host round trips per call
    (25-45 ms of an otherwise 210 ms run_interaction); the QR route is two launches, one tiny SVD and one thin GEMM."""
    if E.shape[0] < E.shape[1]:
        U, S, _ = torch.linalg.svd(E, full_matrices=False)
        keep = S >= EPS_SMALL
        return U[:, keep] * S[keep], None
    if os.environ.get("CRM_CONTEXT_SVD") != "qr":
        # Well-conditioned contexts (the usual case): V from the k x k Gram E'E -- one thin GEMM and a 3 KB read-back instead of the
        # Householder QR (~160 small launches, 3 ms at n = 100k).  U S = E V only needs V orthogonal, which eigh delivers to rounding
        # whatever the conditioning; what the Gram cannot do is tell a singular value below sqrt(eps) from zero (the reference's
        # filter), so anything with sigma_min / sigma_max < 3e-5 takes the QR route below.
        lam, Vg = np.linalg.eigh((E.T @ E).cpu().numpy())
        if lam[0] > 1e-9 * lam[-1] and lam[0] > 1e3 * EPS_SMALL ** 2:
            Vnp = np.empty(Vg.shape, dtype=np.float64, order="C")    # descending singular values, like the SVD (a fresh C-ordered array)
            Vnp[...] = Vg[:, ::-1]
            return E @ torch.from_numpy(Vnp).to(E.device), Vnp
    R = torch.linalg.qr(E, mode="r").R
    _, S, Vh = np.linalg.svd(R.cpu().numpy())
    keep = S >= EPS_SMALL
    Vnp = np.ascontiguousarray(Vh[keep, :].T)
    return E @ torch.from_numpy(Vnp).to(E.device), Vnp


def _L_concat(hK, E, with_map=False):
    """[L_1 | ... | L_k] with L_i = diag(U_i S_i) hK  (reference get_L_values, :533-545), one device tensor; with_map: also the
    k x r matrix V with U S = E V (crm_set_background_factors)."""
    us, V = _scaled_left_vectors(E)
    n = hK.shape[0]
    L = (us[:, :, None] * hK[:, None, :]).reshape(n, us.shape[1] * hK.shape[1]).contiguous()
    return (L, V) if with_map else L


class _LValues(list):
    """The list get_L_values returns, remembering how it was built (hK, E and the map V with U S = E V): a model constructed from it
    with the same contexts can tell the library about the structure (crm_set_background_factors), which checks it against the blocks."""
    factors = None


def get_L_values(hK, E):
    """List of L_i such that K o EE' = sum_i L_i L_i' (reference :533-545); numpy in, numpy out."""
    E = np.asarray(E, float)
    hK = np.asarray(hK, float)
    U, S, Vt = np.linalg.svd(E, full_matrices=False)
    keep = S >= EPS_SMALL
    us = U[:, keep] * S[keep]
    out = _LValues(us[:, i][:, None] * hK for i in range(us.shape[1]))
    if E.shape[0] >= E.shape[1] and hK.ndim == 2:
        out.factors = (hK, E, np.ascontiguousarray(Vt[keep, :].T))
    return out


# element types of genotype matrices understood by the C ABI (include/crm_b200.h: CRM_G_*)
_G_DTYPES = {np.dtype(np.float64): 0, np.dtype(np.int8): 1, np.dtype(np.uint8): 2, np.dtype(np.bool_): 2, np.dtype(np.int16): 3,
             np.dtype(np.int32): 4, np.dtype(np.float32): 5, np.dtype(np.int64): 6}


class _Genotypes:
    """Genotype matrix as handed to the library: pointer, leading dimension (in elements), element type, placement.

    Device tensors are used in place (float64, or int8 dosages; other types are converted on the device).  Host matrices -- numpy
    arrays or CPU tensors, pageable or pinned, float64 like the reference's `asarray(G, float)` or any integer / float32 storage --
    are passed as they are, including column slices of a larger row-major array: the library converts them to int8 dosage blocks with
    its host threads (or moves pinned float64 by DMA) while the device works, see crm_stage_genotypes_typed."""

    def __init__(self, G, device, n, force_float64=False):
        if isinstance(G, torch.Tensor) and G.is_cuda:
            G = G.to(device=device)
            if G.dtype != torch.int8 or force_float64:
                G = G.to(dtype=torch.float64)
            if G.ndim == 2 and (G.stride(1) != 1 or G.stride(0) < G.shape[1]):
                G = G.contiguous()
            assert G.ndim == 2, "G must be n x p"
            self.keep, self.on_host = G, 0
            self.ptr, self.ld = G.data_ptr(), (G.stride(0) if G.shape[0] > 1 else G.shape[1])
            self.dtype = 1 if G.dtype == torch.int8 else 0
        else:
            if isinstance(G, torch.Tensor):
                if G.dtype in (torch.bfloat16, torch.float16) or G.is_complex():      # no numpy view / not a storage the feeder converts
                    G = G.to(torch.float64)
                arr = G.detach().numpy()
            else:
                arr = np.asarray(G)
            assert arr.ndim == 2, "G must be n x p"
            if arr.dtype not in _G_DTYPES or force_float64:
                arr = np.asarray(arr, dtype=np.float64)
            item = arr.dtype.itemsize
            if arr.shape[1] > 0 and arr.shape[0] > 1 and (arr.strides[1] != item or arr.strides[0] % item or arr.strides[0] < arr.shape[1] * item):
                arr = np.ascontiguousarray(arr)
            elif arr.shape[1] > 0 and arr.shape[0] <= 1 and arr.strides[1] != item:
                arr = np.ascontiguousarray(arr)
            self.keep = (G, arr)                  # the caller's object owns the memory
            self.on_host = 1
            self.ptr = arr.ctypes.data
            self.ld = arr.strides[0] // item if arr.shape[0] > 1 else arr.shape[1]
            self.dtype = _G_DTYPES[arr.dtype]
            self.host_array = arr
        shape = self.keep.shape if not self.on_host else self.keep[1].shape
        assert shape[0] == n, "G must have one row per sample"
        self.p = int(shape[1])
        self.rows = int(shape[0])

    def flags(self, donor_level=False):
        return self.on_host | (2 if donor_level else 0) | (self.dtype << 4)

    def rows_permuted(self, idx, device):
        """G[idx, :] for the permuted tested design (float64; a rare path)."""
        if self.on_host:
            return _Genotypes(np.ascontiguousarray(np.asarray(self.host_array, dtype=np.float64)[np.asarray(idx), :]), device, self.rows)
        return _Genotypes(self.keep[torch.as_tensor(np.asarray(idx), device=device), :].to(torch.float64).contiguous(), device, self.rows)


class CellRegMap:
    """Mixed model with genetic-effect heterogeneity across cellular contexts (reference :23-440).

        y = W a + g b1 + g.b2 + e + u + eps,   b2 ~ N(0, v3 E0 E0'),  e ~ N(0, v1 rho1 E1 E1'),
        u ~ N(0, v1 (1-rho1) K o E2 E2'),  eps ~ N(0, v2 I)
    """

    def __init__(self, y, E, W=None, Ls=None, E1=None, hK=None, device=None, _prefetch=None, _group=None, _background_factors=None,
                 _integer_genotypes_likely=False):
        self._device = _device(device)
        dev = self._device
        self._y = _to_dev(y, dev).flatten()
        self._E0 = _to_dev(E, dev)
        n = self._y.shape[0]
        Ls = [] if Ls is None else Ls
        self._W = _to_dev(W, dev) if W is not None else torch.ones((n, 1), dtype=torch.float64, device=dev)
        self._E1 = _to_dev(E1, dev) if E1 is not None else self._E0
        if isinstance(Ls, torch.Tensor):      # pre-concatenated [L_1 | ... | L_k]
            Lcat = _to_dev(Ls, dev)
            n_blocks = 1
        else:
            blocks = [_to_dev(L, dev) for L in Ls]
            for L in blocks:
                assert L.ndim == 2
                assert n == L.shape[0]
            n_blocks = len(blocks)
            Lcat = torch.cat(blocks, dim=1).contiguous() if n_blocks else None
        assert self._W.ndim == 2
        assert self._E0.ndim == 2
        assert self._E1.ndim == 2
        assert n == self._W.shape[0]
        assert n == self._E0.shape[0]
        assert n == self._E1.shape[0]
        self._background_from_Ls = n_blocks > 0
        if n_blocks == 0:
            if hK is None:
                self._rho1 = [1.0]                       # reference :103-106
            else:
                Lcat = _to_dev(hK, dev, two_d=True)      # reference :107-116
                assert Lcat.shape[0] == n
                self._rho1 = np.linspace(0, 1, 11)
        else:
            self._rho1 = np.linspace(0, 1, 11)           # reference :117-131 (hK ignored)
        self._L = Lcat
        # one read-back for all finiteness checks.  y and W: the ValueErrors of glimix_core.lmm.LMM; E, E1, Ls/hK: the reference fails
        # inside LAPACK (SVD of the half-covariance) instead
        checked = [("the outcome", self._y), ("the covariates matrix", self._W), ("E", self._E0), ("E1", self._E1)] + ([("Ls / hK", Lcat)] if Lcat is not None else [])
        finite = torch.stack([torch.isfinite(t).all() for _, t in checked]).cpu().numpy()
        for ok, (name, _) in zip(finite, checked):
            if not ok:
                raise ValueError(f"There are non-finite values in {name}.")
        self._handle = ctypes.c_void_p(0)
        torch.cuda.set_device(dev)
        _lib.call("crm_create", ctypes.byref(self._handle), dev.index if dev.index is not None else torch.cuda.current_device())
        # A host-resident genotype matrix handed over by run_* starts its transfer now, under the set-up.  Every other input is on
        # the device already: a later host-to-device copy would queue behind this transfer on the copy engine.
        self._prefetched = None
        if _prefetch is not None and os.environ.get("CRM_NO_STAGE") != "1" and not (isinstance(_prefetch, torch.Tensor) and _prefetch.is_cuda) \
                and getattr(_prefetch, "ndim", 0) == 2:
            geno = _Genotypes(_prefetch, dev, int(_prefetch.shape[0]))
            if geno.on_host and geno.p > 0:
                width = int(self._E1.shape[1]) + (0 if Lcat is None else int(Lcat.shape[1])) + 1 + int(self._W.shape[1])
                k0 = int(self._E0.shape[1])
                basis_cols = (1 + k0) * (width + (width & 1))
                if _background_factors is not None:      # compact basis of a structured background (layout in abi.cu: kr_apply)
                    q = int(_background_factors[0].shape[1])
                    basis_cols = (width + (width & 1)) + k0 * int(self._E1.shape[1]) + k0 * (1 + int(self._W.shape[1])) + k0 * (k0 + 1) // 2 * q
                _lib.call("crm_stage_genotypes_typed", self._handle, ctypes.c_void_p(geno.ptr), geno.dtype, geno.ld, geno.rows, geno.p, basis_cols, _stream())
            self._prefetched = (_prefetch, geno)
        self._background_factors = None
        if _background_factors is None and isinstance(Ls, _LValues) and Ls.factors is not None and len(Ls) == Ls.factors[2].shape[1]:
            # Ls straight from get_L_values(hK, E) with the contexts of this model (the reference's documented way to build a model)
            hK_l, E_l, V_l = Ls.factors
            if tuple(E_l.shape) == tuple(self._E0.shape) and bool(torch.equal(_to_dev(E_l, dev), self._E0)):
                _background_factors = (_to_dev(hK_l, dev, two_d=True), V_l)
        if _background_factors is not None:
            # Ls = get_L_values(hK, E) with U S = E V: declared ahead of the set-up, which verifies it (crm_set_background_factors)
            hKf, V = _background_factors
            self._background_factors = (hKf, np.ascontiguousarray(V, dtype=np.float64))
            hKf, V = self._background_factors
            _lib.call("crm_set_background_factors", self._handle, _ptr(hKf), hKf.stride(0), int(hKf.shape[1]),
                      V.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(V.shape[1]), None, _stream())
        if _integer_genotypes_likely:
            _lib.call("crm_hint_integer_genotypes", self._handle, 1)
        rho = np.ascontiguousarray(np.asarray(self._rho1, dtype=np.float64))
        mL = 0 if Lcat is None else int(Lcat.shape[1])
        setup_args = (self._handle, _ptr(self._y), _ptr(self._W), self._W.stride(0), _ptr(self._E0),
                      self._E0.stride(0), _ptr(self._E1), self._E1.stride(0), _ptr(Lcat), 0 if Lcat is None else Lcat.stride(0),
                      n, int(self._W.shape[1]), int(self._E0.shape[1]), int(self._E1.shape[1]), mL,
                      rho.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(rho.shape[0]))
        world = 1
        if _group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(None if _group is True else _group)
        if world > 1 and rho.shape[0] > 1:
            self._shared_setup(setup_args, int(rho.shape[0]), None if _group is True else _group)
        else:
            _lib.call("crm_setup", *setup_args, _stream())
        dims = (ctypes.c_int64 * 8)()
        _lib.call("crm_get_dims", self._handle, dims)
        self._dims = {"n": dims[0], "c": dims[1], "k0": dims[2], "m": dims[3], "R": dims[4], "mp": dims[5], "max_rank": dims[6]}

    def _shared_setup(self, setup_args, R, group):
        """Set-up shared between the ranks of `group`: rank r decomposes the grid points r, r + world, ...; one all-gather of the packed
        grid points (S0 and T_rho, ~8 MB each at m = 1 020) over NCCL / NVLink gives every rank the whole basis."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        _lib.call("crm_setup_partial", *setup_args, rank, world, _stream())
        rec = int(_lib.load().crm_basis_record_size(self._handle))
        slots = -(-R // world)                                  # grid points per rank, padded
        mine = torch.zeros((slots, rec), dtype=torch.float64, device=self._device)
        for j, r in enumerate(range(rank, R, world)):
            _lib.call("crm_export_basis", self._handle, r, _ptr(mine[j]), _stream())
        everything = torch.empty((world * slots, rec), dtype=torch.float64, device=self._device)
        dist.all_gather_into_tensor(everything, mine, group=group)
        everything = everything.view(world, slots, rec)
        for other in range(world):
            if other == rank:
                continue
            for j, r in enumerate(range(other, R, world)):
                _lib.call("crm_import_basis", self._handle, r, _ptr(everything[other, j]), _stream())
        _lib.call("crm_setup_finish", self._handle, _stream())

    def _structured_rotation(self):
        """True when the rotation contracts the compact basis of a structured background (crm_rotation_rows)."""
        full, used = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.call("crm_rotation_rows", self._handle, ctypes.byref(full), ctypes.byref(used))
        return used.value < full.value

    def _declare_background_factors(self, hK, V):
        """Tells the library that Ls = get_L_values(hK, E) with U S = E V (crm_set_background_factors); verified there."""
        self._background_factors = (hK, np.ascontiguousarray(V, dtype=np.float64))
        hK, V = self._background_factors
        ok = ctypes.c_int(0)
        _lib.call("crm_set_background_factors", self._handle, _ptr(hK), hK.stride(0), int(hK.shape[1]),
                  V.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(V.shape[1]), ctypes.byref(ok), _stream())
        return bool(ok.value)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.load().crm_destroy(h)
                h.value = None
            except Exception:   # interpreter shutdown
                pass

    @property
    def n_samples(self):
        return int(self._y.shape[0])

    def _pre_expanded_basis(self):
        """True / False once a float64 rotation of this model has chosen its route, None before (tests, bench labels)."""
        dims = (ctypes.c_int64 * 8)()
        _lib.call("crm_get_dims", self._handle, dims)
        return None if dims[7] < 0 else bool(dims[7])

    def set_phenotype(self, y):
        """Extension (not in the reference API): replace the phenotype of this model, keeping cells, contexts, covariates
        and background.  Same results as constructing a new model with the new `y`, without redoing the decompositions
        -- for scans of many genes over one data set."""
        ynew = _to_dev(y, self._device).flatten()
        assert ynew.shape[0] == self.n_samples
        if not bool(torch.isfinite(ynew).all()):
            raise ValueError("There are non-finite values in the outcome.")
        torch.cuda.set_device(self._device)
        _lib.call("crm_update_phenotype", self._handle, _ptr(ynew), _stream())
        self._y = ynew
        return self

    # ------------------------------------------------------------------------------------------
    def _genotypes(self, G, donor_index):
        """(genotype descriptor, flag bits for the C ABI).  With `donor_index` (n,) G is the d x p donor-level matrix
        (G_cells = G[donor_index]); the model then contracts over donors instead of cells (extension, not in the
        reference API)."""
        dev = self._device
        if donor_index is None:
            if self._prefetched is not None and self._prefetched[0] is G:       # staged by the constructor: same host buffer
                geno, self._prefetched = self._prefetched[1], None
                assert geno.rows == self.n_samples, "G must have one row per sample"
            else:
                geno = _Genotypes(G, dev, self.n_samples)
            return geno, geno.flags()
        idx = torch.as_tensor(np.asarray(donor_index.cpu() if isinstance(donor_index, torch.Tensor) else donor_index), device=dev).long().flatten()
        assert idx.numel() == self.n_samples, "donor_index needs one entry per cell"
        d = int(G.shape[0])
        assert idx.numel() == 0 or (int(idx.min()) >= 0 and int(idx.max()) < d), "donor_index out of range"
        key = (d, hash(idx.cpu().numpy().tobytes()))
        if getattr(self, "_donor_key", None) != key:
            perm = torch.argsort(idx, stable=True).to(torch.int32).contiguous()
            counts = torch.bincount(idx, minlength=d)
            offsets = torch.zeros(d + 1, dtype=torch.int64, device=dev)
            offsets[1:] = torch.cumsum(counts, 0)
            offsets = offsets.to(torch.int32).contiguous()
            _lib.call("crm_set_donors", self._handle, _ptr(perm), _ptr(offsets), d, _stream())
            self._donor_key = key
            self._donor_idx = idx
        geno = _Genotypes(G, dev, d)
        return geno, geno.flags(donor_level=True)

    def _expand(self, G, donor_index):
        idx = np.asarray(donor_index.cpu() if isinstance(donor_index, torch.Tensor) else donor_index).ravel()
        if isinstance(G, torch.Tensor):
            return G[torch.as_tensor(idx, device=G.device)]
        return np.asarray(G, float)[idx]

    # ------------------------------------------------------------------------------------------
    def _scan_interaction_device(self, G, idx_E=None, idx_G=None, diagnostics=False, overrides=None, donor_index=None):
        dev = self._device
        torch.cuda.set_device(dev)
        if donor_index is not None and idx_G is not None:   # permuted genotypes break the donor grouping: expand
            G, donor_index = self._expand(G, donor_index), None
        geno, gflags = self._genotypes(G, donor_index)
        p, R, k = geno.p, len(self._rho1), int(self._E0.shape[1])
        out = {name: torch.empty(p, dtype=torch.float64, device=dev) for name in ("pv", "rho1", "e2", "g2", "eps2")}
        diag = _lib.ScanDiag()
        extra = {}
        extra["flags"] = torch.zeros(p, dtype=torch.int32, device=dev)
        diag.flags = extra["flags"].data_ptr()
        if diagnostics:
            extra.update(
                lml=torch.empty((p, R), dtype=torch.float64, device=dev), delta=torch.empty((p, R), dtype=torch.float64, device=dev),
                scale=torch.empty((p, R), dtype=torch.float64, device=dev), Q=torch.empty(p, dtype=torch.float64, device=dev),
                lam=torch.empty((p, k), dtype=torch.float64, device=dev), nlam=torch.empty(p, dtype=torch.int32, device=dev),
                M=torch.empty((p, k, k), dtype=torch.float64, device=dev), liu=torch.empty(p, dtype=torch.float64, device=dev),
                ifault=torch.empty(p, dtype=torch.int32, device=dev), nfev=torch.empty((p, R), dtype=torch.int32, device=dev))
            for name in ("lml", "delta", "scale", "Q", "lam", "nlam", "M", "liu", "ifault", "nfev"):
                setattr(diag, name, extra[name].data_ptr())
        keep = []
        if overrides is not None:
            ridx = torch.as_tensor(np.asarray(overrides["rho_idx"]), dtype=torch.int32, device=dev).contiguous()
            ov0 = _to_dev(overrides["v0"], dev)
            ov1 = _to_dev(overrides["v1"], dev)
            keep += [ridx, ov0, ov1]
            diag.ov_rho_idx, diag.ov_v0, diag.ov_v1 = ridx.data_ptr(), ov0.data_ptr(), ov1.data_ptr()
        gtest = None
        if idx_G is not None:      # row-permuted genotypes in the tested design only (reference :410-413); both designs in float64
            if geno.dtype != 0:
                geno = _Genotypes(geno.host_array if geno.on_host else geno.keep, dev, self.n_samples, force_float64=True)
                gflags = geno.flags()
            gtest = geno.rows_permuted(idx_G, dev)
        if idx_E is not None:      # row-permuted contexts in the tested design only (reference :398-401)
            idx = torch.as_tensor(np.asarray(idx_E), device=dev)
            Etest = self._E0[idx, :].contiguous()
            _lib.call("crm_set_test_contexts", self._handle, _ptr(Etest), Etest.stride(0), _stream())
        if PROFILE["on"]:
            _lib.call("crm_profile", self._handle, 1, None, None, None)
        try:
            _lib.call("crm_scan_interaction", self._handle, ctypes.c_void_p(geno.ptr), geno.ld, p, gflags,
                      ctypes.c_void_p(gtest.ptr if gtest is not None else 0), gtest.ld if gtest is not None else 0, _ptr(out["pv"]), _ptr(out["rho1"]), _ptr(out["e2"]), _ptr(out["g2"]),
                      _ptr(out["eps2"]), ctypes.byref(diag), _stream())
        finally:
            if idx_E is not None:
                _lib.call("crm_set_test_contexts", self._handle, _ptr(self._E0), self._E0.stride(0), _stream())
                if getattr(self, "_background_factors", None) is not None:
                    self._declare_background_factors(*self._background_factors)
        if PROFILE["on"]:
            dims = (ctypes.c_int64 * 8)()
            _lib.call("crm_get_dims", self._handle, dims)
            if dims[7] >= 0:
                PROFILE["pre_expanded_basis"] = bool(dims[7])
            ms, fl, nl = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int64(0)
            _lib.call("crm_profile", self._handle, 0, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(nl))
            PROFILE["rot_ms"] += ms.value
            PROFILE["rot_flops"] += fl.value
            PROFILE["rot_launches"] += nl.value
            full, used = ctypes.c_int64(0), ctypes.c_int64(0)
            _lib.call("crm_rotation_rows", self._handle, ctypes.byref(full), ctypes.byref(used))
            PROFILE["rot_flops_executed"] = PROFILE.get("rot_flops_executed", 0.0) + fl.value * used.value / max(1, full.value)
            PROFILE["rotation_rows"] = (int(full.value), int(used.value))
            _lib.call("crm_profile_int8", self._handle, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(nl))
            PROFILE["int8_ms"] += ms.value
            PROFILE["int8_ops"] += fl.value
            PROFILE["int8_launches"] += nl.value
        out.update(extra)
        return out

    def scan_interaction(self, G, idx_E: Optional[any] = None, idx_G: Optional[any] = None, *, donor_index=None):
        """Score test of H0: v3 = 0 for every column of G (reference :317-440).
        Returns (pvalues (p,), {"rho1", "e2", "g2", "eps2": (p,)}).
        Extension: with `donor_index` (n,), G is the d x p donor-level genotype matrix (G_cells = G[donor_index])."""
        out = self._scan_interaction_device(G, idx_E, idx_G, donor_index=donor_index)
        flags = out["flags"].cpu().numpy()
        if np.any(flags & 1):
            raise RuntimeError("No eigenvalue is bigger than 0!!")
        if np.any(flags & 4):
            raise ValueError("The determinant of H should be positive.")
        info = {key: out[key].cpu().numpy() for key in ("rho1", "e2", "g2", "eps2")}
        return out["pv"].cpu().numpy(), info

    def _scan_association(self, G, fast, donor_index=None):
        dev = self._device
        torch.cuda.set_device(dev)
        geno, gflags = self._genotypes(G, donor_index)
        pv = torch.empty(geno.p, dtype=torch.float64, device=dev)
        alt = torch.empty(geno.p, dtype=torch.float64, device=dev)
        info4 = torch.empty(4, dtype=torch.float64, device=dev)
        null = torch.empty(1, dtype=torch.float64, device=dev)
        _lib.call("crm_scan_association", self._handle, ctypes.c_void_p(geno.ptr), geno.ld, geno.p, gflags,
                  1 if fast else 0, _ptr(pv), _ptr(alt), _ptr(info4), _ptr(null), _stream())
        i4 = info4.cpu().numpy()
        info = {"rho1": i4[0:1].copy(), "e2": i4[1:2].copy(), "g2": i4[2:3].copy(), "eps2": i4[3:4].copy()}
        self._last_association = {"alt_lml": alt, "null_lml": null}
        return pv.cpu().numpy(), info

    def predict_interaction(self, G, MAF, *, donor_index=None):
        """Effect sizes of every column of G (reference :137-205): persistent effect beta_g (p,) and per-cell GxC effects
        beta_gxe (1, n, p).  Like the reference, only the Ls background enters this model (hK given to the constructor
        is not used here) and MAF must lie strictly between 0 and 1."""
        dev = self._device
        torch.cuda.set_device(dev)
        geno, gflags = self._genotypes(G, donor_index)
        p, n = geno.p, self.n_samples
        maf = _to_dev(np.atleast_1d(np.asarray(MAF.detach().cpu().numpy() if isinstance(MAF, torch.Tensor) else MAF, float)), dev)
        assert maf.numel() == p, "one MAF per SNP"
        beta_g = torch.empty(p, dtype=torch.float64, device=dev)
        beta_gxe = torch.empty((n, p), dtype=torch.float64, device=dev)
        rho1 = torch.empty(p, dtype=torch.float64, device=dev)
        _lib.call("crm_predict_interaction", self._handle, ctypes.c_void_p(geno.ptr), geno.ld, p, gflags, _ptr(maf),
                  1 if self._background_from_Ls else 0, _ptr(beta_g), _ptr(beta_gxe), p, _ptr(rho1), _stream())
        self._last_predict = {"rho1": rho1}
        return beta_g.cpu().numpy(), beta_gxe.cpu().numpy().reshape(1, n, p)

    def scan_association(self, G, *, donor_index=None):
        """LRT for a persistent effect of every column of G (reference :246-281)."""
        return self._scan_association(G, fast=False, donor_index=donor_index)

    def scan_association_fast(self, G, *, donor_index=None):
        """Same with delta frozen at the null fit (reference :284-314)."""
        return self._scan_association(G, fast=True, donor_index=donor_index)


def lrt_pvalues(null_lml, alt_lmls, dof=1):
    """Likelihood-ratio p-values (reference :443-469)."""
    dev = _device()
    alt = _to_dev(np.atleast_1d(np.asarray(alt_lmls, float)), dev)
    pv = torch.empty_like(alt)
    _lib.call("crm_lrt_pvalues_dof", _ptr(alt), float(null_lml), alt.numel(), float(dof), _ptr(pv), _stream())
    return pv.cpu().numpy()


def run_association(y, W, E, G, hK=None, *, donor_index=None):
    """Association test (reference :471-500).  NB the reference passes (y, W, E) positionally to
    CellRegMap(y, E, W): W becomes the context/background matrix and E the covariates; kept."""
    crm = CellRegMap(y, W, E, hK=hK)
    return crm.scan_association(G, donor_index=donor_index)


def run_association_fast(y, W, E, G, hK=None, *, donor_index=None):
    """Fast association test (reference :502-531); same positional quirk as run_association."""
    crm = CellRegMap(y, W, E, hK=hK)
    return crm.scan_association_fast(G, donor_index=donor_index)


def _make_interaction_model(y, E, W, E1, E2, hK, device=None, prefetch=None, group=None, donor_level=False):
    dev = _device(device)
    E_dev = _to_dev(E, dev)
    E1 = E_dev if E1 is None else E1
    same_contexts = E2 is None or E2 is E
    E2_dev = E_dev if same_contexts else _to_dev(E2, dev)
    Ls, factors = None, None
    if hK is not None:
        hK_dev = _to_dev(hK, dev, two_d=True)
        Ls, V = _L_concat(hK_dev, E2_dev, with_map=True)
        if V is not None and not same_contexts:
            same_contexts = tuple(E2_dev.shape) == tuple(E_dev.shape) and bool(torch.equal(E2_dev, E_dev))
        if V is not None and same_contexts:
            # the background was built from the tested contexts themselves: L.E_j products are symmetric triple products (see the header)
            factors = (hK_dev, V)
    return CellRegMap(y=y, E=E_dev, W=W, E1=E1, Ls=Ls, device=dev, _prefetch=prefetch, _group=group, _background_factors=factors,
                      _integer_genotypes_likely=not donor_level)


def run_interaction(y, E, G, W=None, E1=None, E2=None, hK=None, idx_G=None, *, donor_index=None):
    """Interaction test (reference :547-587).  NB `idx_G` is forwarded as scan_interaction's second
    positional argument, i.e. it permutes the rows of E (reference :586); kept.
    Extension: with `donor_index` (n,), G is the d x p donor-level genotype matrix (G_cells = G[donor_index])."""
    crm = _make_interaction_model(y, E, W, E1, E2, hK, prefetch=G if donor_index is None else None, donor_level=donor_index is not None)
    return crm.scan_interaction(G, idx_G, donor_index=donor_index)


def estimate_betas(y, W, E, G, maf=None, E1=None, E2=None, hK=None, *, donor_index=None):
    """Effect-size estimator (reference :640-682): returns (beta_g (p,), beta_gxe (1, n, p))."""
    crm = _make_interaction_model(y, E, W, E1, E2, hK, donor_level=True)      # (no rotation of genotypes through digit planes here)
    if maf is None:      # reference: MAF of the expanded matrix
        maf = compute_maf(G if donor_index is None else crm._expand(G, donor_index))
    return crm.predict_interaction(G, maf, donor_index=donor_index)


def compute_maf(X):
    """Minor allele frequencies of dosage columns, NaN = missing (numpy branch of reference :589-638)."""
    if isinstance(X, torch.Tensor):
        X = X.detach().cpu().numpy()
    X = np.asarray(X, float)
    s0 = np.nansum(X, axis=0) / (2 * np.logical_not(np.isnan(X)).sum(axis=0))
    return np.minimum(s0, 1 - s0)
