"""The arithmetic pin: the y vectors of the REFERENCE'S OWN, UNMODIFIED entry points
(spmv/src/dspmv_mgpu_{baseline,v1,v2}.cu, compiled where they lie with oracle/compat_csrmv.h into
oracle/_ref/libref_spmv.so) on a B200.

  not gpu : the oracle (CPU restatement) against the committed vectors tests/golden/ref_y.npz
            (generated on the GPU box by tests/golden/make_golden_y.py)
  gpu     : this repo's library against the same vectors, and against the reference code run
            live on the same host arrays for every visible GPU count

Tolerance (BASELINE.json north_star): per row |y - y_ref| <= 1e-12 * (|alpha| sum|a||x| + |beta||y|).
Rows on a task border of v2 are left out when y != 0 and beta != 0: there the reference keeps ONE
y2 per task and subtracts the wrong original (a test below documents it)."""
import os

import numpy as np
import pytest

import oracle
import ref_cases
from conftest import check_tol


def _golden():
    if not os.path.exists(ref_cases.GOLDEN_Y):
        pytest.fail("tests/golden/ref_y.npz is missing (generate it with tests/golden/make_golden_y.py on a GPU box)")
    return np.load(ref_cases.GOLDEN_Y)


def _rows_to_compare(c, entry, ngpu=1):
    keep = np.ones(c["m"], bool)
    if entry.startswith("v2") and c["beta"] != 0.0 and np.any(c["y0"] != 0.0):
        keep[ref_cases.shared_rows(c, entry, ngpu)] = False
    if entry.startswith("v2"):
        # a limit of the compat layer, not of the reference: cusparseCreateCsr (generic API) refuses a shard with
        # nnz > rows * cols (possible here because rows may repeat a column), the legacy csrmv did not check.
        # The reference's run_task ignores the status and copies the task's y back unchanged: leave those rows out.
        nb, _ = ref_cases.v2_params(c["nnz"], entry)
        p = oracle.generate_tasks_v2(c["rp"], max(nb // ngpu, 1))
        for t in range(len(p["start_idx"])):
            if int(p["dev_nnz"][t]) > int(p["dev_m"][t]) * c["n"]:
                keep[int(p["start_row"][t]):int(p["end_row"][t]) + 1] = False
    return keep


def test_oracle_matches_reference_vectors():
    g = _golden()
    for name, c in ref_cases.cases().items():
        want = oracle.csr_spmv(c["rp"], c["col"], c["val"], c["x"], c["alpha"], c["beta"], c["y0"])
        bound = oracle.csr_spmv_bound(c["rp"], c["col"], c["val"], c["x"], c["alpha"], c["beta"], c["y0"])
        for entry in ref_cases.ENTRIES:
            keep = _rows_to_compare(c, entry)
            check_tol(want[keep], g["%s/%s" % (name, entry)][keep], bound[keep], "oracle vs reference %s/%s" % (name, entry))


def test_oracle_multi_gpu_restatements_match_reference_vectors():
    """The oracle's restatements of the whole entry points (partition + per-shard csrmv + the
    reference's host merge arithmetic) against the reference's vectors, incl. the shared rows."""
    g = _golden()
    for name, c in ref_cases.cases().items():
        bound = oracle.csr_spmv_bound(c["rp"], c["col"], c["val"], c["x"], c["alpha"], c["beta"], c["y0"])
        a = (c["rp"], c["col"], c["val"], c["x"], c["alpha"], c["beta"], c["y0"])
        check_tol(oracle.spmv_mgpu_baseline(*a, 1), g[name + "/baseline"], bound, name + " baseline")
        check_tol(oracle.spmv_mgpu_v1(*a, 1), g[name + "/v1k1"], bound, name + " v1")
        for entry in ("v2k1", "v2k2"):
            nb, _ = ref_cases.v2_params(c["nnz"], entry)
            keep = _rows_to_compare(c, entry)
            got = oracle.spmv_mgpu_v2(*a, nb)
            check_tol(got[keep], g["%s/%s" % (name, entry)][keep], bound[keep], "%s %s" % (name, entry))


def test_reference_v2_single_y2_defect_is_real():
    """With y != 0 and beta != 0 the reference's v2 is off on rows shared by tasks split at both ends
    (one y2 per task, dspmv_mgpu_v2.cu:249,263); its own vectors show it, which is why those rows are
    excluded above and why the library does not reproduce that arithmetic."""
    g = _golden()
    c = ref_cases.cases()["qh768_y"]
    want = oracle.csr_spmv(c["rp"], c["col"], c["val"], c["x"], c["alpha"], c["beta"], c["y0"])
    bound = oracle.csr_spmv_bound(c["rp"], c["col"], c["val"], c["x"], c["alpha"], c["beta"], c["y0"])
    sh = ref_cases.shared_rows(c, "v2k1")
    err = np.abs(g["qh768_y/v2k1"] - want) / bound
    assert (err[np.setdiff1d(np.arange(c["m"]), sh)] <= 1e-12).all()
    assert len(sh) > 0          # informational: the shared rows may or may not be hit, depending on the split
    print("reference v2 error on shared rows (x 1e-12 bound):", (err[sh] / 1e-12).round(2))


# ----------------------------------------------------------------------------- GPU
def _gpu_counts():
    import torch
    n = torch.cuda.device_count()
    want = int(os.environ.get("SBLAS_EXPECT_GPUS", "0"))
    assert n >= want, "SBLAS_EXPECT_GPUS=%d but only %d visible" % (want, n)
    return [g for g in (1, 2, 4, 8) if g <= n]


@pytest.mark.gpu
def test_library_matches_reference_vectors():
    import sblas_b200 as sb
    g = _golden()
    for name, c in ref_cases.cases().items():
        bound = oracle.csr_spmv_bound(c["rp"], c["col"], c["val"], c["x"], c["alpha"], c["beta"], c["y0"])
        for entry in ref_cases.ENTRIES:
            rc, y = ref_cases.run_entry(sb, c, entry, 1)
            assert rc == 0, sb.last_error()
            keep = _rows_to_compare(c, entry)
            check_tol(y[keep], g["%s/%s" % (name, entry)][keep], bound[keep], "library vs reference vectors %s/%s" % (name, entry))


@pytest.mark.gpu
def test_library_matches_reference_code_live():
    """Same host arrays through the reference's own code and through this library, every visible GPU
    count (the reference drives GPUs 0..ngpu-1 from one process, as the library's drop-in entry points do)."""
    import sblas_b200 as sb
    ref = oracle.ref_spmv()
    assert ref is not None, "oracle/_ref/libref_spmv.so was not built"
    for ngpu in _gpu_counts():
        for name, c in ref_cases.cases().items():
            bound = oracle.csr_spmv_bound(c["rp"], c["col"], c["val"], c["x"], c["alpha"], c["beta"], c["y0"])
            for entry in ref_cases.ENTRIES:
                rc_r, y_r = ref_cases.run_entry(ref, c, entry, ngpu)
                if not entry.startswith("v2"):
                    assert rc_r == 0, (name, entry, ngpu)
                rc, y = ref_cases.run_entry(sb, c, entry, ngpu)
                assert rc == 0, sb.last_error()
                keep = _rows_to_compare(c, entry, ngpu)
                # both sides are within 1e-12 of the exact row sum -> 2e-12 between them
                check_tol(y[keep], y_r[keep], 2.0 * bound[keep], "library vs reference code %s/%s ngpu=%d" % (name, entry, ngpu))
