"""s-blas_b200 -- B200-native multi-GPU CSR SpMV behind the s-BLAS API.

Python mirror of the reference interface for ONE path (pnnl/s-blas
spmv/include/spmv_kernel.h:11-36): spMV_mgpu_baseline / spMV_mgpu_v1 / spMV_mgpu_v2
with the same names, argument meaning and return codes, plus the plan API of
include/sblas_spmv.h.  Everything here is a ctypes call into
s-blas_b200/lib/libsblas_spmv.so (host C + hand-written sm_100a kernels); there is
no Python or CPU arithmetic and no fallback: if the library is missing the import
of `lib()` raises.

(The directory name contains '-', so import it as `sblas_b200` via the alias module
at the repo root.)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SBLAS_LIB") or os.path.join(_HERE, "lib", "libsblas_spmv.so")   # SBLAS_LIB: A/B builds

BASELINE, V1, V2, V1_BYTES = 0, 1, 2, 3
ROW_BYTES = 28
SRC_HOST, SRC_DEVICE_SHARD, LAYOUT_ONLY = 0, 1, 2
K_VECTOR, K_TILE, K_TMA, K_VECP = 1, 2, 3, 4
COLS_PREFIX, COLS_BANDED, COLS_UNIFORM, COLS_CIRCUIT, COLS_BANDRUN = 0, 1, 2, 3, 4

_LL = C.c_longlong
_vp = C.c_void_p


class Part(C.Structure):
    """struct sblas_part (include/sblas_spmv.h) == the partition fields of struct spmv_task."""
    _fields_ = [("start_idx", _LL), ("end_idx", _LL), ("start_row", C.c_int), ("end_row", C.c_int),
                ("start_flag", C.c_int), ("end_flag", C.c_int), ("dev_m", C.c_int), ("dev_nnz", C.c_int)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class SegArgs(C.Structure):
    """struct sblas_seg_args (include/sblas_device.h)."""
    _fields_ = [("val", _vp), ("col", _vp), ("rowptr", _vp), ("x", _vp), ("y", _vp), ("edge", _vp),
                ("carry", _vp), ("tail", _vp), ("tstart", _vp), ("tmeta", _vp), ("alpha", C.c_double), ("beta", C.c_double),
                ("row_lo", C.c_int), ("row_hi", C.c_int), ("nz0", C.c_int), ("nz1", C.c_int),
                ("skip_first", C.c_int), ("skip_last", C.c_int), ("tile0", C.c_int), ("ntile", C.c_int),
                ("nz_total", C.c_int), ("mode", C.c_int)]


_lib = None


def lib():
    """Load libsblas_spmv.so.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libsblas_spmv.so is not built (run `make` or __graft_entry__.build()): " + LIB_PATH)
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    mg = [C.c_int, C.c_int, _LL, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int]
    for name, extra in (("baseline", []), ("v1", [C.c_int]), ("v2", [C.c_int, _LL, C.c_int])):
        for pre in ("sblas_spmv_mgpu_", "spMV_mgpu_"):
            f = getattr(L, pre + name)
            f.argtypes = mg + extra
            f.restype = C.c_int
    L.sblas_spmv_cache_clear.restype = None
    L.sblas_get_row_from_index.argtypes = [C.c_int, _vp, _LL]
    L.sblas_get_time.restype = C.c_double
    L.sblas_get_gpu_availble_mem.argtypes = [C.c_int]
    L.sblas_get_gpu_availble_mem.restype = C.c_double
    L.sblas_partition_baseline.argtypes = [C.c_int, _vp, C.c_int, P(Part)]
    L.sblas_partition_v1.argtypes = [C.c_int, _LL, _vp, C.c_int, P(Part)]
    L.sblas_partition_bytes.argtypes = [C.c_int, _LL, _vp, C.c_int, C.c_int, P(Part)]
    L.sblas_v2_num_tasks.argtypes = [_LL, _LL]
    L.sblas_generate_tasks_v2.argtypes = [C.c_int, _LL, _vp, _LL, P(Part)]
    L.sblas_v2_task_owner.argtypes = [C.c_int, C.c_int, C.c_int]
    L.sblas_local_rowptr.argtypes = [_vp, P(Part), C.c_int, _vp]
    L.sblas_local_rowptr.restype = None
    L.sblas_spmv_plan_create.argtypes = [P(_vp), C.c_int, C.c_int, C.c_int, _LL, _vp, _vp, _vp, C.c_int, C.c_int,
                                         _LL, C.c_int]
    L.sblas_spmv_plan_create_rank.argtypes = [P(_vp), C.c_int, C.c_int, C.c_int, _LL, _vp, _vp, _vp, C.c_int,
                                              C.c_int, C.c_int, C.c_int, _LL, C.c_int, C.c_int]
    L.sblas_spmv_plan_execute.argtypes = [_vp, P(C.c_double), _vp, P(C.c_double), _vp]
    L.sblas_spmv_plan_upload.argtypes = [_vp, _vp, _vp]
    L.sblas_spmv_plan_download.argtypes = [_vp, _vp]
    L.sblas_spmv_plan_execute_device.argtypes = [_vp, C.c_double, C.c_double, C.c_int]
    L.sblas_spmv_plan_step.argtypes = [_vp, C.c_double, C.c_double]
    L.sblas_spmv_plan_bind_peer_x.argtypes = [_vp, P(_vp), P(_vp)]
    L.sblas_spmv_plan_num_devices.argtypes = [_vp]
    L.sblas_spmv_plan_num_segments.argtypes = [_vp]
    L.sblas_spmv_plan_segment.argtypes = [_vp, C.c_int, P(Part), P(C.c_int)]
    L.sblas_spmv_plan_x.argtypes = [_vp, C.c_int]
    L.sblas_spmv_plan_x.restype = _vp
    L.sblas_spmv_plan_y.argtypes = [_vp, C.c_int, P(C.c_int), P(C.c_int)]
    L.sblas_spmv_plan_y.restype = _vp
    L.sblas_spmv_plan_rowptr.argtypes = [_vp, C.c_int, P(C.c_int)]
    L.sblas_spmv_plan_rowptr.restype = _vp
    L.sblas_spmv_plan_stream.argtypes = [_vp, C.c_int]
    L.sblas_spmv_plan_stream.restype = _vp
    L.sblas_spmv_plan_edges.argtypes = [_vp, _vp]
    L.sblas_spmv_plan_edge_ptr.argtypes = [_vp, C.c_int]
    L.sblas_spmv_plan_edge_ptr.restype = _vp
    L.sblas_spmv_plan_edge_slots.argtypes = [_vp]
    L.sblas_spmv_plan_merge_gathered.argtypes = [_vp, _vp, C.c_double, C.c_double]
    L.sblas_spmv_plan_bind_edge_table.argtypes = [_vp, _vp]
    L.sblas_spmv_plan_bind_peer_tables.argtypes = [_vp, P(_vp), _LL]
    L.sblas_spmv_plan_exchange_merge.argtypes = [_vp, C.c_double, C.c_double]
    L.sblas_spmv_plan_exchange_merge_phase.argtypes = [_vp, C.c_double, C.c_double, C.c_int]
    L.sblas_spmv_plan_local_segments.argtypes = [_vp]
    L.sblas_spmv_plan_local_segment.argtypes = [_vp, C.c_int, P(_LL)]
    L.sblas_spmv_plan_merge_list.argtypes = [_vp, C.c_int, P(C.c_int), P(P(C.c_int)), P(P(C.c_int)), P(P(_LL))]
    L.sblas_memcpy.argtypes = [_vp, _vp, C.c_ulonglong, C.c_int]
    L.sblas_spmv_plan_alg_bytes.argtypes = [_vp, C.c_int, _LL]
    L.sblas_spmv_plan_alg_bytes.restype = C.c_double
    L.sblas_spmv_plan_launches.argtypes = [_vp]
    L.sblas_spmv_plan_num_units.argtypes = [_vp]
    L.sblas_spmv_plan_x_window.argtypes = [_vp, C.c_int, P(_LL), P(_LL)]
    L.sblas_spmv_plan_chain.argtypes = [_vp]
    L.sblas_mtx_info.argtypes = [C.c_char_p, P(C.c_int), P(C.c_int), P(_LL), P(C.c_int)]
    L.sblas_mtx_read_csr.argtypes = [C.c_char_p, _vp, _vp, _vp]
    L.sblas_spmv_plan_unit.argtypes = [_vp, C.c_int, P(_LL)]
    L.sblas_spmv_plan_execute_unit.argtypes = [_vp, C.c_int, C.c_double, C.c_double]
    L.sblas_spmv_plan_destroy.argtypes = [_vp]
    L.sblas_spmv_plan_destroy.restype = None
    L.sblas_last_error.restype = C.c_char_p
    L.sblas_tile_size.argtypes = [C.c_int]
    L.sblas_launch_rebase_rowptr.argtypes = [_vp, _LL, C.c_int, _LL, _vp, _vp]
    L.sblas_launch_tile_rows.argtypes = [P(SegArgs), C.c_int, _vp, _vp]
    L.sblas_launch_tile_meta.argtypes = [P(SegArgs), C.c_int, _vp, _vp]
    L.sblas_launch_row_block_stats.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]
    L.sblas_tile_size_kind.argtypes = [C.c_int, C.c_int]
    L.sblas_launch_spmv_segment.argtypes = [P(SegArgs), C.c_int, C.c_int, C.c_int, _vp]
    L.sblas_launch_edge_merge.argtypes = [_vp, _vp, _vp, C.c_int, _vp, C.c_double, C.c_double, _vp]
    L.sblas_launch_fill_f64.argtypes = [_vp, _LL, C.c_double, _vp]
    L.sblas_synth_fill_csr.argtypes = [_vp, C.c_int, C.c_int, _LL, _LL, C.c_int, C.c_int, _LL, C.c_ulonglong,
                                       C.c_int, C.c_double, _vp, _vp, _vp]
    L.sblas_synth_fill_uniform.argtypes = [_vp, _LL, C.c_ulonglong, C.c_double, C.c_double, _vp]
    sm = [C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int]
    for name in ("sblas_spmm_mgpu", "cusparse_mgpu_csrmm", "cusparse_mgpu_csrmm_omp"):
        getattr(L, name).argtypes = sm
    L.sblas_spmm_plan_create.argtypes = [P(_vp), C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int]
    L.sblas_spmm_plan_execute.argtypes = [_vp, C.c_int, P(C.c_double), _vp, P(C.c_double), _vp]
    L.sblas_spmm_plan_execute_device.argtypes = [_vp, C.c_int, C.c_int, C.c_double, _vp, C.c_double, _vp, C.c_int]
    L.sblas_spmm_plan_columns.argtypes = [_vp, C.c_int, C.c_int, P(C.c_int), P(C.c_int)]
    L.sblas_spmm_plan_stream.argtypes = [_vp, C.c_int]
    L.sblas_spmm_plan_stream.restype = _vp
    L.sblas_spmm_plan_num_devices.argtypes = [_vp]
    L.sblas_spmm_plan_destroy.argtypes = [_vp]
    L.sblas_spmm_plan_destroy.restype = None
    L.sblas_sptrans_mgpu.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]
    L.kernal_sptrans.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.sblas_sptrans_last_device_ms.restype = C.c_double
    L.sblas_synth_read_probe.argtypes = [_vp, C.c_ulonglong, C.c_int, _vp]
    L.sblas_synth_read_probe.restype = C.c_double
    _lib = L
    return L


def last_error():
    return (lib().sblas_last_error() or b"").decode()


# ----------------------------------------------------------------------------- helpers
def _host(a, dt, name):
    a = np.asarray(a)
    if a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("%s must be a C-contiguous numpy array of %s (like the reference's host buffers)" % (name, dt))
    return a


def _ptr(a):
    return a.ctypes.data if isinstance(a, np.ndarray) else int(a)


# ----------------------------------------------------------------------------- reference entry points
def spMV_mgpu_baseline(m, n, nnz, alpha, csrVal, csrRowPtr, csrColIndex, x, beta, y, ngpu):
    """spmv/include/spmv_kernel.h:11-15.  y (numpy float64) is updated in place; returns the status code."""
    a, b = C.c_double(alpha), C.c_double(beta)
    return lib().spMV_mgpu_baseline(m, n, nnz, C.addressof(a), _ptr(_host(csrVal, np.float64, "csrVal")),
                                    _ptr(_host(csrRowPtr, np.int64, "csrRowPtr")),
                                    _ptr(_host(csrColIndex, np.int32, "csrColIndex")),
                                    _ptr(_host(x, np.float64, "x")), C.addressof(b),
                                    _ptr(_host(y, np.float64, "y")), ngpu)


def spMV_mgpu_v1(m, n, nnz, alpha, csrVal, csrRowPtr, csrColIndex, x, beta, y, ngpu, kernel):
    """spmv/include/spmv_kernel.h:16-21."""
    a, b = C.c_double(alpha), C.c_double(beta)
    return lib().spMV_mgpu_v1(m, n, nnz, C.addressof(a), _ptr(_host(csrVal, np.float64, "csrVal")),
                              _ptr(_host(csrRowPtr, np.int64, "csrRowPtr")),
                              _ptr(_host(csrColIndex, np.int32, "csrColIndex")),
                              _ptr(_host(x, np.float64, "x")), C.addressof(b),
                              _ptr(_host(y, np.float64, "y")), ngpu, kernel)


def spMV_mgpu_v2(m, n, nnz, alpha, csrVal, csrRowPtr, csrColIndex, x, beta, y, ngpu, kernel, nb, copy_of_workspace):
    """spmv/include/spmv_kernel.h:23-30."""
    a, b = C.c_double(alpha), C.c_double(beta)
    return lib().spMV_mgpu_v2(m, n, nnz, C.addressof(a), _ptr(_host(csrVal, np.float64, "csrVal")),
                              _ptr(_host(csrRowPtr, np.int64, "csrRowPtr")),
                              _ptr(_host(csrColIndex, np.int32, "csrColIndex")),
                              _ptr(_host(x, np.float64, "x")), C.addressof(b),
                              _ptr(_host(y, np.float64, "y")), ngpu, kernel, int(nb), copy_of_workspace)


def cache_clear():
    """Drop the plans cached by the one-shot entry points (SBLAS_PLAN_CACHE=1)."""
    lib().sblas_spmv_cache_clear()


def get_row_from_index(n, a, idx):
    """spmv/include/spmv_kernel.h:32."""
    return lib().sblas_get_row_from_index(n, _ptr(_host(a, np.int64, "a")), int(idx))


def get_time():
    return lib().sblas_get_time()


def get_gpu_availble_mem(ngpu):
    return lib().sblas_get_gpu_availble_mem(ngpu)


# ----------------------------------------------------------------------------- partitioners
def _parts_to_dict(arr, count):
    keys = [k for k, _ in Part._fields_]
    out = {k: np.zeros(count, np.int64 if k.endswith("idx") else np.int32) for k in keys}
    for i in range(count):
        for k in keys:
            out[k][i] = getattr(arr[i], k)
    return out


def partition_baseline(csrRowPtr, ngpu):
    rp = _host(csrRowPtr, np.int64, "csrRowPtr")
    arr = (Part * ngpu)()
    rc = lib().sblas_partition_baseline(len(rp) - 1, _ptr(rp), ngpu, arr)
    assert rc == 0
    return _parts_to_dict(arr, ngpu)


def partition_v1(csrRowPtr, ngpu):
    rp = _host(csrRowPtr, np.int64, "csrRowPtr")
    arr = (Part * ngpu)()
    rc = lib().sblas_partition_v1(len(rp) - 1, int(rp[-1]), _ptr(rp), ngpu, arr)
    assert rc == 0
    return _parts_to_dict(arr, ngpu)


def partition_bytes(csrRowPtr, ngpu, row_bytes=ROW_BYTES):
    """Opt-in byte-balanced variant of the v1 partition (not in the reference)."""
    rp = _host(csrRowPtr, np.int64, "csrRowPtr")
    arr = (Part * ngpu)()
    rc = lib().sblas_partition_bytes(len(rp) - 1, int(rp[-1]), _ptr(rp), ngpu, row_bytes, arr)
    assert rc == 0
    return _parts_to_dict(arr, ngpu)


def generate_tasks_v2(csrRowPtr, nb):
    rp = _host(csrRowPtr, np.int64, "csrRowPtr")
    T = lib().sblas_v2_num_tasks(int(rp[-1]), int(nb))
    arr = (Part * max(T, 1))()
    got = lib().sblas_generate_tasks_v2(len(rp) - 1, int(rp[-1]), _ptr(rp), int(nb), arr)
    assert got == T
    return _parts_to_dict(arr, T)


def v2_task_owner(T, ngpu, task):
    return lib().sblas_v2_task_owner(T, ngpu, task)


def local_rowptr(csrRowPtr, part, baseline=False):
    rp = _host(csrRowPtr, np.int64, "csrRowPtr")
    p = Part(**{k: int(part[k]) for k, _ in Part._fields_})
    out = np.zeros(p.dev_m + 1, np.int32)
    lib().sblas_local_rowptr(_ptr(rp), C.byref(p), 1 if baseline else 0, _ptr(out))
    return out


# ----------------------------------------------------------------------------- plan
class Plan:
    """Resident multi-GPU SpMV plan (include/sblas_spmv.h, plan API)."""

    def __init__(self, handle, m, n, keep=()):
        self._h = handle
        self.m, self.n = m, n
        self._keep = keep          # arrays that must outlive the plan (adopted device shards)

    @classmethod
    def create(cls, version, m, n, nnz, csrVal, csrRowPtr, csrColIndex, ngpu, kernel=1, nb=0, q=1):
        h = _vp()
        rc = lib().sblas_spmv_plan_create(C.byref(h), version, m, n, nnz,
                                          _ptr(_host(csrVal, np.float64, "csrVal")),
                                          _ptr(_host(csrRowPtr, np.int64, "csrRowPtr")),
                                          _ptr(_host(csrColIndex, np.int32, "csrColIndex")), ngpu, kernel, int(nb), q)
        if rc != 0:
            raise RuntimeError("sblas_spmv_plan_create rc=%d: %s" % (rc, last_error()))
        return cls(h, m, n)

    @classmethod
    def create_rank(cls, version, m, n, nnz, csrVal, csrRowPtr, csrColIndex, world, rank, device, kernel=1,
                    nb=0, q=1, flags=SRC_HOST, keep=()):
        """csrVal / csrColIndex: numpy host arrays (SRC_HOST) or integer device pointers (SRC_DEVICE_SHARD)."""
        h = _vp()
        rc = lib().sblas_spmv_plan_create_rank(C.byref(h), version, m, n, nnz, _ptr(csrVal),
                                               _ptr(_host(csrRowPtr, np.int64, "csrRowPtr")), _ptr(csrColIndex),
                                               world, rank, device, kernel, int(nb), q, flags)
        if rc != 0:
            raise RuntimeError("sblas_spmv_plan_create_rank rc=%d: %s" % (rc, last_error()))
        return cls(h, m, n, keep)

    def execute(self, alpha, x, beta, y):
        a, b = C.c_double(alpha), C.c_double(beta)
        rc = lib().sblas_spmv_plan_execute(self._h, C.byref(a), _ptr(x), C.byref(b), _ptr(y))
        if rc != 0:
            raise RuntimeError("sblas_spmv_plan_execute rc=%d: %s" % (rc, last_error()))

    def upload(self, x, y=None):
        rc = lib().sblas_spmv_plan_upload(self._h, _ptr(x), None if y is None else _ptr(y))
        if rc != 0:
            raise RuntimeError("sblas_spmv_plan_upload rc=%d: %s" % (rc, last_error()))

    def download(self, y):
        rc = lib().sblas_spmv_plan_download(self._h, None if y is None else _ptr(y))
        if rc != 0:
            raise RuntimeError("sblas_spmv_plan_download rc=%d: %s" % (rc, last_error()))

    def execute_device(self, alpha, beta, sync=False):
        rc = lib().sblas_spmv_plan_execute_device(self._h, alpha, beta, 1 if sync else 0)
        if rc != 0:
            raise RuntimeError("sblas_spmv_plan_execute_device rc=%d: %s" % (rc, last_error()))

    def step(self, alpha, beta):
        """One product with x and y resident (kernels + fused split-row exchange of a bound rank plan); replayed
        from a CUDA graph on single-GPU plans.  Asynchronous on the plan's stream."""
        rc = lib().sblas_spmv_plan_step(self._h, alpha, beta)
        if rc != 0:
            raise RuntimeError("sblas_spmv_plan_step rc=%d: %s" % (rc, last_error()))

    def x_window(self, dev=0):
        """[first, last] column the GPU's shard references (what upload() copies of x)."""
        a, b = _LL(), _LL()
        assert lib().sblas_spmv_plan_x_window(self._h, dev, C.byref(a), C.byref(b)) == 0
        return int(a.value), int(b.value)

    def chain(self):
        """x <- y on every GPU of the plan (device-side all-gather over NVLink); then execute_device."""
        rc = lib().sblas_spmv_plan_chain(self._h)
        if rc != 0:
            raise RuntimeError("chain failed: %s" % last_error())

    def bind_peer_x(self, peer_x_ptrs, peer_flag_ptrs):
        """Rank plans: compute on the peer-mapped x buffer peer_x_ptrs[rank]; chain() then all-gathers y into every
        rank's x over NVLink (peer_flag_ptrs[r]: rank r's 2*world zeroed 8-byte words)."""
        a = (_vp * len(peer_x_ptrs))(*[int(p) for p in peer_x_ptrs])
        b = (_vp * len(peer_flag_ptrs))(*[int(p) for p in peer_flag_ptrs])
        rc = lib().sblas_spmv_plan_bind_peer_x(self._h, a, b)
        if rc != 0:
            raise RuntimeError("sblas_spmv_plan_bind_peer_x rc=%d: %s" % (rc, last_error()))

    def merge_gathered(self, gathered_ptr, alpha, beta):
        rc = lib().sblas_spmv_plan_merge_gathered(self._h, int(gathered_ptr), alpha, beta)
        if rc != 0:
            raise RuntimeError("sblas_spmv_plan_merge_gathered rc=%d: %s" % (rc, last_error()))

    @property
    def num_devices(self):
        return lib().sblas_spmv_plan_num_devices(self._h)

    @property
    def num_segments(self):
        return lib().sblas_spmv_plan_num_segments(self._h)

    def segment(self, i):
        p, d = Part(), C.c_int()
        assert lib().sblas_spmv_plan_segment(self._h, i, C.byref(p), C.byref(d)) == 0
        out = p.as_dict()
        out["device"] = d.value
        return out

    def x_ptr(self, dev=0):
        return lib().sblas_spmv_plan_x(self._h, dev)

    def y_ptr(self, dev=0):
        fr, rows = C.c_int(), C.c_int()
        p = lib().sblas_spmv_plan_y(self._h, dev, C.byref(fr), C.byref(rows))
        return p, fr.value, rows.value

    def rowptr_ptr(self, dev=0):
        cnt = C.c_int()
        p = lib().sblas_spmv_plan_rowptr(self._h, dev, C.byref(cnt))
        return p, cnt.value

    def stream(self, dev=0):
        return lib().sblas_spmv_plan_stream(self._h, dev)

    def edge_ptr(self, dev=0):
        return lib().sblas_spmv_plan_edge_ptr(self._h, dev)

    def local_segments(self):
        """Host layout: the segments this process runs (see sblas_spmv_plan_local_segment)."""
        keys = ("gidx", "row_lo", "row_hi", "nz0", "nz1", "shared_first", "shared_last", "edge_slot", "dev", "dev_first_row")
        out = []
        for i in range(lib().sblas_spmv_plan_local_segments(self._h)):
            buf = (_LL * 10)()
            assert lib().sblas_spmv_plan_local_segment(self._h, i, buf) == 0
            out.append(dict(zip(keys, [int(v) for v in buf])))
        return out

    def merge_list(self, dev=0):
        """Host layout: [(GPU-local row, [offsets into the rank-major edge table])] this GPU finishes."""
        n = C.c_int()
        mrow, mbeg, moff = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(_LL)()
        assert lib().sblas_spmv_plan_merge_list(self._h, dev, C.byref(n), C.byref(mrow), C.byref(mbeg), C.byref(moff)) == 0
        return [(mrow[i], [int(moff[k]) for k in range(mbeg[i], mbeg[i + 1])]) for i in range(n.value)]

    def bind_peer_tables(self, peer_ptrs, table_words):
        """peer_ptrs[r] = rank r's exchange buffer as mapped here (2*table_words + 2*world words, zeroed)."""
        arr = (_vp * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        rc = lib().sblas_spmv_plan_bind_peer_tables(self._h, arr, int(table_words))
        if rc != 0:
            raise RuntimeError("sblas_spmv_plan_bind_peer_tables rc=%d: %s" % (rc, last_error()))

    def exchange_merge(self, alpha, beta, phase=0):
        rc = lib().sblas_spmv_plan_exchange_merge_phase(self._h, alpha, beta, phase)
        if rc != 0:
            raise RuntimeError("sblas_spmv_plan_exchange_merge rc=%d: %s" % (rc, last_error()))

    def bind_edge_table(self, device_ptr):
        assert lib().sblas_spmv_plan_bind_edge_table(self._h, int(device_ptr)) == 0

    @property
    def edge_slots(self):
        return lib().sblas_spmv_plan_edge_slots(self._h)

    def edges(self):
        out = np.zeros(2 * self.num_segments, np.float64)
        assert lib().sblas_spmv_plan_edges(self._h, _ptr(out)) == 0
        return out

    def alg_bytes(self, beta_nonzero, x_touched_per_gpu=-1):
        return lib().sblas_spmv_plan_alg_bytes(self._h, 1 if beta_nonzero else 0, int(x_touched_per_gpu))

    @property
    def launches(self):
        return lib().sblas_spmv_plan_launches(self._h)

    def units(self):
        """Row panels of the plan (one kernel each): list of dicts."""
        out = []
        for i in range(lib().sblas_spmv_plan_num_units(self._h)):
            b = (_LL * 8)()
            assert lib().sblas_spmv_plan_unit(self._h, i, b) == 0
            out.append(dict(index=i, segment=int(b[0]), kind=int(b[1]), ipt=int(b[2]), row_lo=int(b[3]), row_hi=int(b[4]),
                            nz0=int(b[5]), nz1=int(b[6]), launches=int(b[7])))
        return out

    def execute_unit(self, i, alpha, beta):
        rc = lib().sblas_spmv_plan_execute_unit(self._h, int(i), float(alpha), float(beta))
        if rc != 0:
            raise RuntimeError("execute_unit failed: %s" % last_error())

    def destroy(self):
        if self._h:
            lib().sblas_spmv_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def mtx_read_csr(path):
    """Correct Matrix-Market -> CSR ingest (include/sblas_ingest.h; opt-in):
    returns (m, n, rowptr int64, col int32, val float64, is_symmetric)."""
    m, n, nnz, sym = C.c_int(), C.c_int(), _LL(), C.c_int()
    rc = lib().sblas_mtx_info(path.encode(), C.byref(m), C.byref(n), C.byref(nnz), C.byref(sym))
    if rc != 0:
        raise IOError("sblas_mtx_info(%s) = %d" % (path, rc))
    rp = np.zeros(m.value + 1, np.int64)
    col = np.zeros(max(nnz.value, 1), np.int32)[:nnz.value]
    val = np.zeros(max(nnz.value, 1), np.float64)[:nnz.value]
    rc = lib().sblas_mtx_read_csr(path.encode(), rp.ctypes.data, col.ctypes.data if nnz.value else None,
                                  val.ctypes.data if nnz.value else None)
    if rc != 0:
        raise IOError("sblas_mtx_read_csr(%s) = %d" % (path, rc))
    return m.value, n.value, rp, col, val, bool(sym.value)


def memcpy(dst, src, nbytes, kind):
    """kind: 1 H2D, 2 D2H, 3 D2D (blocking).  dst/src: numpy arrays or integer device pointers."""
    rc = lib().sblas_memcpy(_ptr(dst), _ptr(src), int(nbytes), kind)
    if rc != 0:
        raise RuntimeError("sblas_memcpy: cuda error %d" % rc)


def device_synchronize():
    rc = lib().sblas_device_synchronize()
    if rc != 0:
        raise RuntimeError("cudaDeviceSynchronize: cuda error %d" % rc)


# ----------------------------------------------------------------------------- synthetic content (device)
def synth_fill_csr(d_rowptr, row_first, nrows, k0, k1, n, cols_mode, band, seed, d_val, d_col, value_const=None,
                   stream=None):
    rc = lib().sblas_synth_fill_csr(int(d_rowptr), row_first, nrows, int(k0), int(k1), n, cols_mode, int(band),
                                    int(seed), 0 if value_const is None else 1,
                                    0.0 if value_const is None else float(value_const), int(d_val), int(d_col),
                                    stream)
    if rc != 0:
        raise RuntimeError("sblas_synth_fill_csr: cuda error %d" % rc)


def synth_fill_uniform(d_p, count, seed, lo=0.0, hi=1.0, stream=None):
    rc = lib().sblas_synth_fill_uniform(int(d_p), int(count), int(seed), lo, hi, stream)
    if rc != 0:
        raise RuntimeError("sblas_synth_fill_uniform: cuda error %d" % rc)


def synth_read_probe(d_buf, nbytes, reps=3, stream=None):
    """Read-only HBM bandwidth (GB/s) of a streaming-load kernel over nbytes of device memory."""
    return lib().sblas_synth_read_probe(int(d_buf), int(nbytes), int(reps), stream)


# ----------------------------------------------------------------------------- SpMM
def _colmajor(a, rows, cols, name):
    """B / C as the reference holds them: column-major.  Accepts a Fortran-ordered (rows, cols) array or a flat
    array of rows*cols doubles; returns the flat view (no copy)."""
    a = np.asarray(a)
    if a.dtype != np.float64:
        raise TypeError(name + " must be float64")
    if a.ndim == 2:
        if a.shape != (rows, cols) or not a.flags["F_CONTIGUOUS"]:
            raise TypeError("%s must be a Fortran-ordered (%d, %d) array (column-major, like the reference's buffers)" % (name, rows, cols))
        return a.reshape(-1, order="F")
    if a.size != rows * cols or not a.flags["C_CONTIGUOUS"]:
        raise TypeError(name + " must hold rows*cols contiguous doubles")
    return a


def cusparse_mgpu_csrmm(m, n, k, alpha, nnz_A, csrRowPtr_A, csrColIndex_A, csrVal_A, beta, B_dense, C_dense, ngpu, omp=False):
    """spmm/include/spmm_kernel.h:6-31 (omp=True: the _omp entry point).  C_dense is updated in place."""
    a, b = C.c_double(alpha), C.c_double(beta)
    fn = lib().cusparse_mgpu_csrmm_omp if omp else lib().cusparse_mgpu_csrmm
    Bf, Cf = _colmajor(B_dense, k, n, "B_dense"), _colmajor(C_dense, m, n, "C_dense")
    return fn(m, n, k, C.addressof(a), nnz_A, _ptr(_host(csrRowPtr_A, np.int32, "csrRowPtr_A")),
              _ptr(_host(csrColIndex_A, np.int32, "csrColIndex_A")), _ptr(_host(csrVal_A, np.float64, "csrVal_A")),
              C.addressof(b), _ptr(Bf), _ptr(Cf), ngpu)


class SpmmPlan:
    """A resident on every GPU of the plan (include/sblas_spmm.h)."""

    def __init__(self, m, k, nnz, rowptr32, col, val, ngpu):
        h = _vp()
        rc = lib().sblas_spmm_plan_create(C.byref(h), m, k, nnz, _ptr(_host(rowptr32, np.int32, "csrRowPtr_A")),
                                          _ptr(_host(col, np.int32, "csrColIndex_A")), _ptr(_host(val, np.float64, "csrVal_A")), ngpu)
        if rc != 0:
            raise RuntimeError("sblas_spmm_plan_create rc=%d: %s" % (rc, last_error()))
        self._h, self.m, self.k, self.ngpu = h, m, k, ngpu

    def execute(self, n, alpha, B, beta, Cm):
        a, b = C.c_double(alpha), C.c_double(beta)
        rc = lib().sblas_spmm_plan_execute(self._h, n, C.byref(a), _ptr(_colmajor(B, self.k, n, "B")), C.byref(b),
                                           _ptr(_colmajor(Cm, self.m, n, "C")))
        if rc != 0:
            raise RuntimeError("sblas_spmm_plan_execute rc=%d: %s" % (rc, last_error()))

    def execute_device(self, dev, nd, alpha, d_B, beta, d_C, sync=False):
        rc = lib().sblas_spmm_plan_execute_device(self._h, dev, nd, alpha, int(d_B), beta, int(d_C), 1 if sync else 0)
        if rc != 0:
            raise RuntimeError("sblas_spmm_plan_execute_device rc=%d: %s" % (rc, last_error()))

    def columns(self, n, dev):
        a, b = C.c_int(), C.c_int()
        assert lib().sblas_spmm_plan_columns(self._h, n, dev, C.byref(a), C.byref(b)) == 0
        return a.value, b.value

    def stream(self, dev=0):
        return lib().sblas_spmm_plan_stream(self._h, dev)

    def destroy(self):
        if self._h:
            lib().sblas_spmm_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


# ----------------------------------------------------------------------------- transposition
def kernal_sptrans(m, n, nnz, ngpu, csrRowPtr, csrColIdx, csrVal, ref=None):
    """CSR -> CSC on ngpu GPUs (sptrans/sptrans_v1/src/sptrans_kernal.h:80, the reference's entry point).
    Returns (status, cscColPtr, cscRowIdx, cscVal); ref = (colptr, rowidx, val) host arrays to be compared like the
    reference's *_ref arguments (status 2 on a mismatch)."""
    rp, cc, vv = _host(csrRowPtr, np.int32, "csrRowPtr"), _host(csrColIdx, np.int32, "csrColIdx"), _host(csrVal, np.float64, "csrVal")
    colptr, rowidx, val = np.zeros(n + 1, np.int32), np.zeros(max(nnz, 1), np.int32), np.zeros(max(nnz, 1), np.float64)
    r = [None, None, None] if ref is None else [_ptr(_host(ref[1], np.int32, "ref rowidx")), _ptr(_host(ref[0], np.int32, "ref colptr")),
                                                _ptr(_host(ref[2], np.float64, "ref val"))]
    rc = lib().kernal_sptrans(m, n, nnz, ngpu, _ptr(rp), _ptr(cc), _ptr(vv), _ptr(rowidx), _ptr(colptr), _ptr(val), *r)
    return rc, colptr, rowidx[:nnz], val[:nnz]


def sptrans_last_device_ms():
    return lib().sblas_sptrans_last_device_ms()
