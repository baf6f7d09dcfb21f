"""oracle -- TEST INFRASTRUCTURE ONLY (the checker, never the product).

ctypes front-end of oracle/spmv_oracle.c (the CPU restatement of the s-BLAS
multi-GPU CSR SpMV path; see that file's header for the parity status and the
reference file:line each function follows) and of oracle/_ref/libref_helper.so
(the reference's own spmv/src/spmv_helper.cu compiled where it lies by
oracle/Makefile) and oracle/_ref/libref_spmv.so (the reference's own, unmodified
three entry points over a cusparseDcsrmv -> cusparseSpMV compat header; GPU only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LL = C.c_longlong
_pd = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_pi = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_pl = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the checker (liboracle.so, and _ref/ when the reference checkout exists)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "spmv_oracle.c")
    stale = force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
    need_ref = os.path.exists("/root/reference/spmv/src/spmv_helper.cu") and not os.path.exists(
        os.path.join(_HERE, "_ref", "libref_helper.so"))
    if stale or need_ref:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.oracle_csr_spmv.argtypes = [C.c_int, _pl, _pi, _pd, _pd, C.c_double, C.c_double, _pd]
        L.oracle_csr_spmv_omp.argtypes = L.oracle_csr_spmv.argtypes
        L.oracle_csr_spmv_omp.restype = C.c_int
        L.oracle_csr_spmv_omp_balanced.argtypes = L.oracle_csr_spmv.argtypes
        L.oracle_csr_spmv_omp_balanced.restype = C.c_int
        L.oracle_csr_spmv_bound.argtypes = [C.c_int, _pl, _pi, _pd, _pd, C.c_double, C.c_double, _pd, _pd]
        L.oracle_get_row_from_index.argtypes = [C.c_int, _pl, _LL]
        L.oracle_get_row_from_index.restype = C.c_int
        L.oracle_partition_baseline.argtypes = [C.c_int, _pl, C.c_int, _pi, _pi, _pi, _pi]
        L.oracle_local_rowptr_baseline.argtypes = [_pl, C.c_int, C.c_int, _pi]
        L.oracle_partition_v1.argtypes = [C.c_int, _LL, _pl, C.c_int, _pl, _pl, _pi, _pi, _pi, _pi, _pi, _pi]
        L.oracle_local_rowptr_v1.argtypes = [_pl, _LL, C.c_int, C.c_int, C.c_int, _pi]
        L.oracle_v2_num_tasks.argtypes = [_LL, _LL]
        L.oracle_v2_num_tasks.restype = C.c_int
        L.oracle_generate_tasks_v2.argtypes = [C.c_int, _LL, _pl, _LL, _pl, _pl, _pi, _pi, _pi, _pi, _pi, _pi]
        L.oracle_v2_quota.argtypes = [C.c_int, C.c_int, C.c_int]
        L.oracle_v2_quota.restype = C.c_int
        mg = [C.c_int, C.c_int, _LL, C.c_double, _pd, _pl, _pi, _pd, C.c_double, _pd]
        L.oracle_spmv_mgpu_v1.argtypes = mg + [C.c_int]
        L.oracle_spmv_mgpu_baseline.argtypes = mg + [C.c_int]
        L.oracle_spmv_mgpu_v2.argtypes = mg + [_LL]
        L.oracle_spmv_mgpu_v2_ex.argtypes = mg + [_LL, C.c_int]
        L.oracle_gen_g.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_gen_g.restype = _LL
        L.oracle_srand.argtypes = [C.c_uint]
        L.oracle_rand_unit.restype = C.c_double
        L.oracle_coo_to_rowptr.argtypes = [C.c_int, _LL, _pi, _pl]
        L.oracle_load_mtx.argtypes = [C.c_char_p, C.c_char, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_load_mtx.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_set_threads.restype = C.c_int
        L.oracle_csr_check.argtypes = [C.c_int, _pl, _pi, _pd, _pd, C.c_double, C.c_double, _pd, _pd, C.c_int, C.c_int,
                                       _pd, C.POINTER(C.c_int)]
        L.oracle_csr_check.restype = C.c_double
        L.oracle_synth_fill_csr.argtypes = [_pl, C.c_int, C.c_int, _LL, _LL, C.c_int, C.c_int, _LL, C.c_ulonglong,
                                            C.c_int, C.c_double, _pd, _pi]
        L.oracle_synth_fill_csr.restype = None
        L.oracle_synth_fill_uniform.argtypes = [_pd, _LL, C.c_ulonglong, C.c_double, C.c_double]
        L.oracle_synth_fill_uniform.restype = None
        L.oracle_csrmm.argtypes = [C.c_int, C.c_int, _pi, _pi, _pd, _pd, _LL, _pd, _LL, C.c_double, C.c_double]
        L.oracle_csrmm.restype = None
        L.oracle_csrmm_bound.argtypes = [C.c_int, C.c_int, _pi, _pi, _pd, _pd, _LL, _pd, _LL, C.c_double, C.c_double, _pd]
        L.oracle_csrmm_bound.restype = None
        L.oracle_spmm_mgpu.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, _pi, _pi, _pd, C.c_double, _pd, _pd, C.c_int]
        L.oracle_csr2csc.argtypes = [C.c_int, C.c_int, C.c_int, _pi, _pi, _pd, _pi, _pi, _pd]
        L.oracle_csr2csc.restype = None
        _lib = L
    return _lib


def host_cores():
    """Cores this process may run on (the affinity mask, not a launcher's OMP_NUM_THREADS)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def set_threads(n=None):
    """Set the OpenMP thread count of the CPU legs explicitly (default: every core of the affinity
    mask) -- torchrun exports OMP_NUM_THREADS=1 to its ranks.  Returns the count in force."""
    return lib().oracle_set_threads(int(n) if n else host_cores())


def csr_check(local_rowptr, col, val, x, alpha, beta, y_in, y_got, skip_first=-1, skip_last=-1):
    """Full-vector check of one shard on all cores.  Returns (worst |err|/bound over the rows that are not
    skipped, its row, edge = [sum_first, bound_first, sum_last, bound_last] raw partials of the skipped rows)."""
    edge = np.zeros(4)
    wr = C.c_int(-1)
    w = lib().oracle_csr_check(len(local_rowptr) - 1, _c(local_rowptr, np.int64), _c(col, np.int32), _c(val, np.float64),
                               _c(x, np.float64), alpha, beta, _c(y_in, np.float64), _c(y_got, np.float64),
                               int(skip_first), int(skip_last), edge, C.byref(wr))
    return float(w), int(wr.value), edge


def synth_fill_csr(rowptr_slice, row_first, k0, k1, n, cols_mode, band, seed, value_const=None):
    """Host twin of sblas_synth_fill_csr: val/col of the global entry range [k0, k1) of the rows
    [row_first, row_first + len(rowptr_slice) - 1) whose int64 row pointer slice is given."""
    rp = _c(rowptr_slice, np.int64)
    val = np.empty(k1 - k0, np.float64)
    col = np.empty(k1 - k0, np.int32)
    lib().oracle_synth_fill_csr(rp, int(row_first), len(rp) - 1, int(k0), int(k1), int(n), int(cols_mode), int(band),
                                int(seed), 0 if value_const is None else 1, 0.0 if value_const is None else float(value_const),
                                val, col)
    return val, col


def synth_fill_uniform(count, seed, lo=0.0, hi=1.0):
    out = np.empty(count, np.float64)
    lib().oracle_synth_fill_uniform(out, int(count), int(seed), lo, hi)
    return out


def csrmm(rowptr32, col, val, B, alpha, beta, C, ngpu=None):
    """C_out = alpha*A*B + beta*C (oracle_csrmm), B (k x n) and C (m x n) column-major, i.e. numpy arrays of
    shape (n, k) / (n, m) in C order or Fortran-ordered (k, n) / (m, n).  ngpu: go through the column split of the
    reference's entry point (oracle_spmm_mgpu).  Returns a new array shaped like C."""
    rp = _c(rowptr32, np.int32)
    m = len(rp) - 1
    Bf, Cf = np.asfortranarray(B, dtype=np.float64), np.asfortranarray(C, dtype=np.float64).copy(order="F")
    k, n = Bf.shape
    assert Cf.shape == (m, n)
    bflat, cflat = Bf.reshape(-1, order="F"), Cf.reshape(-1, order="F")
    if ngpu is None:
        lib().oracle_csrmm(m, n, rp, _c(col, np.int32), _c(val, np.float64), bflat, k, cflat, m, alpha, beta)
    else:
        assert lib().oracle_spmm_mgpu(m, n, k, alpha, rp, _c(col, np.int32), _c(val, np.float64), beta, bflat, cflat, ngpu) == 0
    return cflat.reshape((m, n), order="F")


def csrmm_bound(rowptr32, col, val, B, alpha, beta, C):
    rp = _c(rowptr32, np.int32)
    m = len(rp) - 1
    Bf, Cf = np.asfortranarray(B, dtype=np.float64), np.asfortranarray(C, dtype=np.float64)
    k, n = Bf.shape
    out = np.empty(m * n)
    lib().oracle_csrmm_bound(m, n, rp, _c(col, np.int32), _c(val, np.float64), Bf.reshape(-1, order="F"), k,
                             Cf.reshape(-1, order="F"), m, alpha, beta, out)
    return out.reshape((m, n), order="F")


def csr2csc(m, n, rowptr32, col, val):
    """CSR -> CSC like the reference's host transposition (oracle_csr2csc).  Returns (colptr, rowidx, val)."""
    rp, cc, vv = _c(rowptr32, np.int32), _c(col, np.int32), _c(val, np.float64)
    nnz = int(rp[-1])
    colptr, rowidx, out = np.zeros(n + 1, np.int32), np.zeros(max(nnz, 1), np.int32), np.zeros(max(nnz, 1), np.float64)
    lib().oracle_csr2csc(m, n, nnz, rp, cc if nnz else np.zeros(1, np.int32), vv if nnz else np.zeros(1), rowidx, colptr, out)
    return colptr, rowidx[:nnz], out[:nnz]


def ref_sptrans():
    """The reference's OWN host transposition (sptrans/sptrans_v1/src/tranpose.h, compiled where it lies behind
    oracle/ref_sptrans_shim.cpp into oracle/_ref/libref_sptrans.so), or None.  Pure host code: runs without a GPU.
    Returns f(m, n, rowptr32, col, val) -> (colptr, rowidx, val)."""
    p = os.path.join(_HERE, "_ref", "libref_sptrans.so")
    if not os.path.exists(p):
        return None
    R = C.CDLL(p)
    R.ref_matrix_transposition.argtypes = [C.c_int, C.c_int, C.c_int, _pi, _pi, _pd, _pi, _pi, _pd]
    R.ref_matrix_transposition.restype = None

    def call(m, n, rowptr32, col, val):
        rp, cc, vv = _c(rowptr32, np.int32), _c(col, np.int32), _c(val, np.float64)
        nnz = int(rp[-1])
        colptr, rowidx, out = np.zeros(n + 1, np.int32), np.zeros(max(nnz, 1), np.int32), np.zeros(max(nnz, 1), np.float64)
        R.ref_matrix_transposition(m, n, nnz, rp, cc if nnz else np.zeros(1, np.int32), vv if nnz else np.zeros(1), rowidx, colptr, out)
        return colptr, rowidx[:nnz], out[:nnz]
    return call


_ref_spmm = None


def ref_spmm():
    """The reference's OWN SpMM entry points (spmm/src/dspmm_mgpu_baseline.cu, unmodified, compiled where it
    lies with oracle/compat_csrmv.h mapping cusparseDcsrmm onto cusparseSpMM) from
    oracle/_ref/libref_spmm.so, or None.  GPU only.  Returns f(m, n, k, alpha, nnz, rowptr32, col, val, beta,
    B_flat, C_flat, ngpu, omp) -> status; C_flat (column-major) is updated in place."""
    global _ref_spmm
    if _ref_spmm is not None:
        return _ref_spmm
    p = os.path.join(_HERE, "_ref", "libref_spmm.so")
    if not os.path.exists(p):
        return None
    R = C.CDLL(p)
    vp = C.c_void_p
    at = [C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int]
    f = getattr(R, "_Z19cusparse_mgpu_csrmmiiiPKdiPiS1_PdS0_S2_S2_i")
    fo = getattr(R, "_Z23cusparse_mgpu_csrmm_ompiiiPKdiPiS1_PdS0_S2_S2_i")
    f.argtypes = at
    fo.argtypes = at

    def call(m, n, k, alpha, nnz, rp, col, val, beta, B, Cm, ngpu, omp=True):
        a, b = C.c_double(alpha), C.c_double(beta)
        return (fo if omp else f)(m, n, k, C.addressof(a), nnz, rp.ctypes.data, col.ctypes.data, val.ctypes.data,
                                  C.addressof(b), B.ctypes.data, Cm.ctypes.data, ngpu)
    _ref_spmm = call
    return call


_ref = None


def ref_helper():
    """The reference's own compiled spmv_helper.cu (get_row_from_index), or None."""
    global _ref
    if _ref is None:
        p = os.path.join(_HERE, "_ref", "libref_helper.so")
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        f = getattr(R, "_Z18get_row_from_indexiPxx")     # C++ mangled, SURVEY.md F9
        f.argtypes = [C.c_int, _pl, _LL]
        f.restype = C.c_int
        _ref = f
    return _ref


_ref_spmv = None


def ref_spmv():
    """The reference's OWN three entry points -- spmv/src/dspmv_mgpu_{baseline,v1,v2}.cu + spmv_helper.cu,
    unmodified, compiled where they lie with oracle/compat_csrmv.h force-included (legacy
    cusparseDcsrmv[_mp] -> cusparseSpMV) into oracle/_ref/libref_spmv.so -- or None when the library
    was not built.  Needs a GPU to run.  Returns an object with baseline / v1 / v2 taking the
    reference's argument list (numpy host arrays, y updated in place) and returning its status code
    (v2's is undefined in the reference: it falls off the end)."""
    global _ref_spmv
    if _ref_spmv is not None:
        return _ref_spmv
    p = os.path.join(_HERE, "_ref", "libref_spmv.so")
    if not os.path.exists(p):
        return None
    R = C.CDLL(p)                       # RTLD_LOCAL + -Bsymbolic: its mangled names never meet the product's
    vp = C.c_void_p
    mg = [C.c_int, C.c_int, _LL, vp, vp, vp, vp, vp, vp, vp, C.c_int]
    fb = getattr(R, "_Z18spMV_mgpu_baselineiixPdS_PxPiS_S_S_i")
    f1 = getattr(R, "_Z12spMV_mgpu_v1iixPdS_PxPiS_S_S_ii")
    f2 = getattr(R, "_Z12spMV_mgpu_v2iixPdS_PxPiS_S_S_iixi")
    fb.argtypes = mg
    f1.argtypes = mg + [C.c_int]
    f2.argtypes = mg + [C.c_int, _LL, C.c_int]

    class Ref:
        path = p

        @staticmethod
        def _call(fn, m, n, nnz, alpha, val, rp, col, x, beta, y, *extra):
            a, b = C.c_double(alpha), C.c_double(beta)
            for arr, dt in ((val, np.float64), (rp, np.int64), (col, np.int32), (x, np.float64), (y, np.float64)):
                assert arr.dtype == dt and arr.flags["C_CONTIGUOUS"]
            return fn(m, n, nnz, C.addressof(a), val.ctypes.data, rp.ctypes.data, col.ctypes.data, x.ctypes.data,
                      C.addressof(b), y.ctypes.data, *extra)

        def baseline(self, m, n, nnz, alpha, val, rp, col, x, beta, y, ngpu):
            return self._call(fb, m, n, nnz, alpha, val, rp, col, x, beta, y, ngpu)

        def v1(self, m, n, nnz, alpha, val, rp, col, x, beta, y, ngpu, kernel):
            return self._call(f1, m, n, nnz, alpha, val, rp, col, x, beta, y, ngpu, kernel)

        def v2(self, m, n, nnz, alpha, val, rp, col, x, beta, y, ngpu, kernel, nb, q):
            return self._call(f2, m, n, nnz, alpha, val, rp, col, x, beta, y, ngpu, kernel, int(nb), q)

    _ref_spmv = Ref()
    return _ref_spmv


# ----------------------------------------------------------------------------- wrappers
def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def csr_spmv(rowptr, col, val, x, alpha, beta, y):
    """y_out = alpha*A*x + beta*y (oracle_csr_spmv). Returns a new array."""
    out = _c(y, np.float64).copy()
    lib().oracle_csr_spmv(len(rowptr) - 1, _c(rowptr, np.int64), _c(col, np.int32), _c(val, np.float64),
                          _c(x, np.float64), alpha, beta, out)
    return out


def csr_spmv_bound(rowptr, col, val, x, alpha, beta, y):
    b = np.empty(len(rowptr) - 1, np.float64)
    lib().oracle_csr_spmv_bound(len(rowptr) - 1, _c(rowptr, np.int64), _c(col, np.int32), _c(val, np.float64),
                                _c(x, np.float64), alpha, beta, _c(y, np.float64), b)
    return b


def get_row_from_index(rowptr, idx):
    rp = _c(rowptr, np.int64)
    return lib().oracle_get_row_from_index(len(rp) - 1, rp, int(idx))


def partition_v1(rowptr, ngpu):
    rp = _c(rowptr, np.int64)
    m, nnz = len(rp) - 1, int(rp[-1])
    si, ei = np.zeros(ngpu, np.int64), np.zeros(ngpu, np.int64)
    a = [np.zeros(ngpu, np.int32) for _ in range(6)]
    lib().oracle_partition_v1(m, nnz, rp, ngpu, si, ei, *a)
    return dict(start_idx=si, end_idx=ei, start_row=a[0], end_row=a[1], start_flag=a[2], end_flag=a[3],
                dev_m=a[4], dev_nnz=a[5])


def partition_baseline(rowptr, ngpu):
    rp = _c(rowptr, np.int64)
    a = [np.zeros(ngpu, np.int32) for _ in range(4)]
    lib().oracle_partition_baseline(len(rp) - 1, rp, ngpu, *a)
    return dict(start_row=a[0], end_row=a[1], dev_m=a[2], dev_nnz=a[3])


def generate_tasks_v2(rowptr, nb):
    rp = _c(rowptr, np.int64)
    m, nnz = len(rp) - 1, int(rp[-1])
    T = lib().oracle_v2_num_tasks(nnz, nb)
    si, ei = np.zeros(T, np.int64), np.zeros(T, np.int64)
    a = [np.zeros(T, np.int32) for _ in range(6)]
    lib().oracle_generate_tasks_v2(m, nnz, rp, nb, si, ei, *a)
    return dict(start_idx=si, end_idx=ei, start_row=a[0], end_row=a[1], start_flag=a[2], end_flag=a[3],
                dev_m=a[4], dev_nnz=a[5])


def local_rowptr_v1(rowptr, start_idx, start_row, dev_m, dev_nnz):
    out = np.zeros(dev_m + 1, np.int32)
    lib().oracle_local_rowptr_v1(_c(rowptr, np.int64), int(start_idx), int(start_row), int(dev_m), int(dev_nnz), out)
    return out


def local_rowptr_baseline(rowptr, start_row, dev_m):
    out = np.zeros(dev_m + 1, np.int32)
    lib().oracle_local_rowptr_baseline(_c(rowptr, np.int64), int(start_row), int(dev_m), out)
    return out


def _mgpu(fn, rowptr, col, val, x, alpha, beta, y, *last):
    rp = _c(rowptr, np.int64)
    out = _c(y, np.float64).copy()
    xx = _c(x, np.float64)
    rc = fn(len(rp) - 1, len(xx), int(rp[-1]), alpha, _c(val, np.float64), rp,
            _c(col, np.int32), xx, beta, out, *last)
    assert rc == 0
    return out


def spmv_mgpu_v1(rowptr, col, val, x, alpha, beta, y, ngpu):
    return _mgpu(lib().oracle_spmv_mgpu_v1, rowptr, col, val, x, alpha, beta, y, ngpu)


def spmv_mgpu_baseline(rowptr, col, val, x, alpha, beta, y, ngpu):
    return _mgpu(lib().oracle_spmv_mgpu_baseline, rowptr, col, val, x, alpha, beta, y, ngpu)


def spmv_mgpu_v2(rowptr, col, val, x, alpha, beta, y, nb, faithful_y2=False):
    """faithful_y2=True reproduces the reference's single-y2 defect (see spmv_oracle.c)."""
    return _mgpu(lib().oracle_spmv_mgpu_v2_ex, rowptr, col, val, x, alpha, beta, y, int(nb), 1 if faithful_y2 else 0)


def gen_g(n, r1=0.9, r2=0.01, seed=1):
    """The harness `g` generator (glibc rand(), srand(seed); the harness never seeds => seed 1).
    Returns (coo_row, coo_col, val, alpha, beta) with ALPHA/BETA drawn right after, as in
    dspmv_test.cu:281-282."""
    L = lib()
    nnz = L.oracle_gen_g(n, r1, r2, None, None, None)
    if nnz < 0:
        raise ValueError("n must be a positive multiple of 8")
    r, c, v = np.empty(nnz, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
    L.oracle_srand(seed)
    L.oracle_gen_g(n, r1, r2, r.ctypes.data, c.ctypes.data, v.ctypes.data)
    alpha = L.oracle_rand_unit()
    beta = L.oracle_rand_unit()
    return r, c, v, alpha, beta


def coo_to_rowptr(m, coo_row):
    rp = np.zeros(m + 1, np.int64)
    lib().oracle_coo_to_rowptr(m, len(coo_row), _c(coo_row, np.int32), rp)
    return rp


def load_mtx(path, mode="f"):
    """The harness loader (file order kept, not row sorted, symmetric flag ignored)."""
    L = lib()
    m, n, nz = C.c_int(), C.c_int(), C.c_int()
    rc = L.oracle_load_mtx(path.encode(), mode.encode(), C.byref(m), C.byref(n), C.byref(nz), None, None, None)
    if rc:
        raise IOError("oracle_load_mtx rc=%d" % rc)
    r, c, v = np.zeros(nz.value, np.int32), np.zeros(nz.value, np.int32), np.zeros(nz.value, np.float64)
    L.oracle_load_mtx(path.encode(), mode.encode(), C.byref(m), C.byref(n), C.byref(nz),
                      r.ctypes.data, c.ctypes.data, v.ctypes.data)
    return m.value, n.value, r, c, v


def load_mtx_csr(path):
    """CPU restatement (numpy) of the reference's CORRECT loader, sptrsv/sptrsv_v1/src/
    mmio_highlevel.h:139-298 (mmio_data): entries bucketed by row in file order, off-diagonal
    entries of symmetric / hermitian files mirrored right after their original with the same value,
    pattern -> 1.0, integer converted, complex -> real part.  Checker for include/sblas_ingest.h
    (SURVEY.md section 8f-1).  Returns (m, n, rowptr int64, col int32, val float64, is_symmetric)."""
    with open(path) as f:
        banner = f.readline().split()
        if len(banner) != 5 or not banner[0].startswith("%%MatrixMarket"):
            raise IOError("bad banner")
        field, symm = banner[3].lower(), banner[4].lower()
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        m, n, listed = (int(t) for t in line.split()[:3])
        sym = symm in ("symmetric", "hermitian")            # mmio_highlevel.h:174-178
        rows, cols, vals = [], [], []
        toks = f.read().split()
    per = {"pattern": 2, "complex": 4}.get(field, 3)
    for e in range(listed):
        t = toks[per * e: per * e + per]
        i, j = int(t[0]) - 1, int(t[1]) - 1
        v = 1.0 if field == "pattern" else float(t[2])
        rows.append(i); cols.append(j); vals.append(v)
        if sym and i != j:                                  # mirrored entry follows its original (:254-266)
            rows.append(j); cols.append(i); vals.append(v)
    rows = np.asarray(rows, np.int64)
    order = np.argsort(rows, kind="stable")                 # bucket by row, file order kept inside a row
    rp = np.zeros(m + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=m), out=rp[1:])
    return m, n, rp, np.asarray(cols, np.int32)[order], np.asarray(vals, np.float64)[order], sym


def ref_ingest():
    """The reference's OWN correct loader (mmio_highlevel.h) compiled by oracle/Makefile, or None."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_ingest.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref_mmio_info.argtypes = [C.POINTER(C.c_int)] * 4 + [C.c_char_p]
    L.ref_mmio_data.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p]

    def load(p):
        m, n, nnz, sym = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rc = L.ref_mmio_info(C.byref(m), C.byref(n), C.byref(nnz), C.byref(sym), p.encode())
        if rc != 0:
            raise IOError("ref mmio_info rc=%d" % rc)
        rp = np.zeros(m.value + 1, np.int32)
        col = np.zeros(max(nnz.value, 1), np.int32)
        val = np.zeros(max(nnz.value, 1), np.float64)
        rc = L.ref_mmio_data(rp.ctypes.data, col.ctypes.data, val.ctypes.data, p.encode())
        if rc != 0:
            raise IOError("ref mmio_data rc=%d" % rc)
        return m.value, n.value, rp.astype(np.int64), col[:nnz.value], val[:nnz.value], bool(sym.value)
    return load
