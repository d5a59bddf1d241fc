"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly the
symbols include/rcb200.h declares; host-only entry points work; compute entry points fail loudly
(never fall back) without a B200; the host mirror reproduces the reference's argument errors."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import esn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "reservoircomputing.jl_b200")


@pytest.fixture(scope="session")
def built():
    sys.path.insert(0, PKG)
    import build as rcb_build

    return rcb_build.build()


@pytest.fixture(scope="session")
def R(built):
    import rcb200

    return rcb200


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "rcb200.h")).read()
    return sorted(set(re.findall(r"RCB_API\s+[\w\s\*]+?\b(rcb_\w+)\s*\(", hdr)))


def test_header_symbols_all_exported(built):
    declared = _declared_symbols()
    assert len(declared) >= 28
    out = subprocess.check_output(["nm", "-D", "--defined-only", built], text=True)
    exported = sorted(set(re.findall(r"\sT\s(rcb_\w+)", out)))
    assert exported == declared, (set(declared) ^ set(exported))


def test_ctypes_signatures_cover_header(R):
    from rcb200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    assert R.lib().rcb_version() == 100


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, "include", "rcb200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)  # declarations only, comments stripped
    assert "torch" not in code and "at::" not in code and "Tensor" not in code
    assert re.findall(r"#include\s+<([^>]+)>", code) == ["stdint.h"]


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".jl")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.replace("cross-check oracle", ""), f"{f} references oracle/"


def test_compute_entry_points_fail_loudly_without_gpu(R):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(R.CudaError, match="no CPU fallback"):
        R.Device(0)
    esn = R.ESN(3, 20, 3)
    ps, st = R.setup(np.random.default_rng(0), esn)
    with pytest.raises(R.CudaError):
        R.collectstates(esn, np.zeros((3, 4), np.float32), ps, st)
    with pytest.raises(R.CudaError):
        R.fit_readout(R.RidgeRegression(1e-3), np.ones((4, 20)), np.ones((2, 20)))


# ---- host-only format conversion: index parity is bit-exact ---------------------------------------
@pytest.mark.parametrize("n,k", [(300, 6), (1024, 6), (64, 1)])
def test_csr_from_dense_bit_exact(R, n, k):
    W = O.rand_sparse_like(np.random.default_rng(17), n, radius=0.9, sparsity=k / n)
    rp, ci, v = R.csr_from_matrix(W)
    rp0, ci0, v0 = O.dense_to_csr(W)
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)
    assert v.tobytes() == v0.astype(np.float32).tobytes()
    assert rp[-1] == n * k


def test_csr_from_csc_matches_dense_path(R):
    Wc = O.rand_sparse_like(np.random.default_rng(3), 200, radius=1.1, sparsity=0.05, as_csc=True)
    a = R.csr_from_matrix(Wc)
    b = R.csr_from_matrix(Wc.toarray())
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()


def test_csr_edge_cases(R):
    Z = np.zeros((5, 5), np.float32)
    rp, ci, v = R.csr_from_matrix(Z)  # empty pattern
    assert rp.tolist() == [0] * 6 and ci.size == 0
    D = np.full((3, 3), 2.0, np.float32)  # fully dense
    rp, ci, v = R.csr_from_matrix(D)
    assert rp.tolist() == [0, 3, 6, 9] and ci.tolist() == [0, 1, 2] * 3
    Wn = np.zeros((3, 3), np.float32); Wn[1, 2] = np.nan
    with pytest.raises(R.NumericError, match="invalid values"):
        R.csr_from_matrix(Wn)
    Wz = np.zeros((2, 2), np.float32); Wz[0, 1] = -0.0  # -0.0 is a structural zero (iszero)
    assert R.csr_from_matrix(Wz)[0][-1] == 0


# ---- host-mirror argument errors mirror the reference's ------------------------------------------
def test_constructor_errors(R):
    with pytest.raises(R.DimensionMismatch, match="leak_coefficient must have length out_dims=6"):
        R.ESNCell(4, 6, leak_coefficient=np.ones(5))  # esn_cell.jl:135-142
    with pytest.raises(R.DimensionMismatch):
        R.DeepESN(3, [10, 10, 10], 3, leak_coefficient=(0.1, 0.2))  # esn_generics.jl:51-66


def test_setup_structure_and_draw_order(R):
    esn = R.ESN(3, 50, 2, init_reservoir=R.rand_sparse(radius=1.2, sparsity=6 / 50), state_modifiers=(R.NLAT2, R.Pad(1.0)))
    ps1, st1 = R.setup(np.random.default_rng(5), esn)
    ps2, st2 = R.setup(np.random.default_rng(5), esn)
    assert ps1 == ps2  # same seed => same parameters (serialization_reproducibility_tests.jl:59-81)
    assert ps1.reservoir.input_matrix.shape == (50, 3) and ps1.reservoir.reservoir_matrix.shape == (50, 50)
    assert "bias" not in ps1.reservoir and ps1.readout.weight.shape == (2, 50)
    assert st1.reservoir.carry is None and "rng" in st1.reservoir.cell
    assert np.all((ps1.reservoir.reservoir_matrix != 0).sum(axis=0) == 6)
    assert abs(O.spectral_radius(ps1.reservoir.reservoir_matrix) - 1.2) < 1e-4
    deep = R.DeepESN(3, [20, 30], 2, state_modifiers=(R.NLAT2,))
    psd, std = R.setup(np.random.default_rng(1), deep)
    assert psd.cells[1].input_matrix.shape == (30, 20) and len(std.cells) == 2


# ---- sibling cells: host-side pieces ----------------------------------------------------------------
@pytest.mark.parametrize("n,k,gamma", [(300, 6, 0.2), (64, 3, 0.0), (128, 128, 1.0)])
def test_eusn_operator_csr_bit_exact(R, n, k, gamma):  # eusn_cell.jl:120-123, entry by entry in Float32
    W = O.rand_sparse_like(np.random.default_rng(23), n, radius=0.9, sparsity=k / n)
    if gamma == 0.0:
        W[5, 9] = W[9, 5] = 0.25  # a symmetric pair cancels to a structural zero
    rp, ci, v = R.eusn_operator_csr(W, gamma)
    A = O.compute_asym_recurrent(W, np.float32(gamma))
    rp0, ci0, v0 = O.dense_to_csr(A)
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)
    assert v.tobytes() == v0.astype(np.float32).tobytes()


def test_sibling_cell_setup_structure(R):  # es2n_cell.jl:93-104: draw order input, reservoir, orthogonal, bias
    rc = R.ES2N(3, 20, 2, proximity=0.4, use_bias=True)
    ps, st = R.setup(np.random.default_rng(4), rc)
    assert list(ps.reservoir.keys()) == ["input_matrix", "reservoir_matrix", "orthogonal_matrix", "bias"]
    Om = ps.reservoir.orthogonal_matrix
    assert Om.shape == (20, 20) and np.allclose(Om.T @ Om, np.eye(20), atol=1e-5)
    rc = R.ResESN(3, 20, 2, alpha=0.8, beta=0.6)
    ps, _ = R.setup(np.random.default_rng(4), rc)
    assert "orthogonal_matrix" in ps.reservoir and "bias" not in ps.reservoir
    rc = R.EuSN(3, 20, 2, diffusion=0.3, leak_coefficient=0.5)
    ps, _ = R.setup(np.random.default_rng(4), rc)
    assert "orthogonal_matrix" not in ps.reservoir and rc.reservoir.cell_params == (0.3, 0.0)


@pytest.mark.parametrize("m,seed", [(1, 0), (2, 1), (5, 2), (40, 3), (200, 4)])
def test_hessenberg_eigvals_match_lapack(R, m, seed):  # host part of rcb_spectral_radius_* 
    rng = np.random.default_rng(seed)
    H = np.triu(rng.standard_normal((m, m)), -1)
    ev = R.hessenberg_eigvals(H)
    ref = np.linalg.eigvals(H)
    key = lambda z: (round(z.real, 8), round(z.imag, 8))
    a, b = np.array(sorted(ev, key=key)), np.array(sorted(ref, key=key))
    assert np.max(np.abs(np.sort(np.abs(ev)) - np.sort(np.abs(ref)))) <= 1e-9 * max(1.0, np.max(np.abs(ref)))
    assert abs(np.sum(ev) - np.trace(H)) <= 1e-9 * m * max(1.0, np.max(np.abs(ref)))
    assert np.max(np.abs(a - b)) <= 1e-7 * max(1.0, np.max(np.abs(ref)))


def test_hessenberg_eigvals_special_cases(R):
    assert np.allclose(np.sort(R.hessenberg_eigvals(np.diag([3.0, -1.0, 2.0])).real), [-1, 2, 3])
    J = np.array([[0.0, -1.0], [1.0, 0.0]])               # a conjugate pair on the unit circle
    ev = R.hessenberg_eigvals(J)
    assert np.allclose(sorted(ev.imag), [-1, 1]) and np.allclose(ev.real, 0)
    C3 = np.array([[0.0, 0, 1], [1, 0, 0], [0, 1, 0]])     # cyclic shift: cube roots of unity (all moduli equal)
    assert np.allclose(np.abs(R.hessenberg_eigvals(C3)), 1.0)


def test_hessenberg_eigvals_both_iterations_agree(R, monkeypatch):
    """The real Francis double-shift iteration (default) and the complex single-shift fallback see the same spectrum."""
    rng = np.random.default_rng(11)
    H = np.triu(rng.standard_normal((120, 120)), -1)
    a = R.hessenberg_eigvals(H)
    monkeypatch.setenv("RCB_HESS_QR", "complex")
    b = R.hessenberg_eigvals(H)
    assert np.max(np.abs(np.sort(np.abs(a)) - np.sort(np.abs(b)))) <= 1e-9 * np.max(np.abs(a))
    assert abs(a.sum() - b.sum()) <= 1e-9 * 120
    assert np.sum(a.imag > 0) == np.sum(a.imag < 0)   # complex eigenvalues of a real matrix come in conjugate pairs


# ---- the kernels' gather schedule is built on the host: its invariants are checkable without a GPU -----------------
@pytest.mark.parametrize("n,k,nthreads", [(64, 3, 64), (300, 6, 320), (1024, 6, 1024), (2048, 12, 1024), (4096, 6, 1024)])
def test_gather_schedule_invariants(R, n, k, nthreads):
    import ctypes as C
    from rcb200 import _lib as L

    W = O.rand_sparse_like(np.random.default_rng(n + k), n, radius=0.9, sparsity=k / n) if n <= 1500 else None
    if W is None:  # no spectral scaling needed for a structural test
        rng = np.random.default_rng(n + k)
        W = np.zeros((n, n), np.float32)
        for j in range(n):
            W[rng.choice(n, size=k, replace=False), j] = rng.standard_normal(k).astype(np.float32)
    Wf = np.asfortranarray(W, dtype=np.float32)
    st = (C.c_int64 * 10)()
    for balanced in (0, 1):
        L.check(L.lib().rcb_sell_stats_from_dense(L.ptr(Wf), n, n, nthreads, balanced, st))
        nnz, padded, wf64, ideal64, wf32, ideal32, roundtrip, missing, wb, wbi = list(st)
        assert nnz == n * k and roundtrip == 1 and missing == 0        # every (row, col, value) triple exactly once, rows a bijection
        assert padded >= nnz and padded % 32 == 0
        if balanced:
            assert wf64 == ideal64                                       # every half-warp LDS.64 gather is bank-conflict free
            if n >= 300:
                assert padded <= 1.35 * nnz                              # the price: a bounded amount of padding ...
            if n >= 4096 and k == 6:
                assert padded <= 1.10 * nnz                              # ... 4-7 % at the sizes the batched kernel is built for
            assert wb <= 1.5 * wbi + 8                                   # rows of a half-warp mostly distinct mod 16 (write-back conflicts)
        else:
            assert wf64 >= ideal64


def test_tile_members_layout():
    """Host logic of the ensemble sweep: a shared series becomes one column-major (d, T) block per member."""
    from rcb200.api import _tile_members

    a = np.asfortranarray(np.arange(3 * 7, dtype=np.float32).reshape(3, 7, 1))
    out = _tile_members(a, 5)
    assert out.shape == (3, 7, 5) and out.flags.f_contiguous and out.dtype == np.float32
    assert np.array_equal(out, np.repeat(a, 5, axis=2))
    assert np.array_equal(np.frombuffer(out.tobytes(order="A"), np.float32)[:21], a.ravel(order="F"))

