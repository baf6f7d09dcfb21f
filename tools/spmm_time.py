"""SpMM timing on one GPU (diagnostic): resident A, device B / C, CUDA events on the plan's stream.
usage: python tools/spmm_time.py [workload=inproc] [n=128] [reps=10]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import oracle
    import sblas_b200 as sb
    wl = sys.argv[1] if len(sys.argv) > 1 else "inproc"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    c = bench.host_problem(wl)
    m, k, nnz = c["m"], c["n"], c["nnz"]
    rp32 = c["rp"].astype(np.int32)
    t0 = time.perf_counter()
    p = sb.SpmmPlan(m, k, nnz, rp32, c["col"], c["val"], 1)
    t_plan = time.perf_counter() - t0
    B = torch.rand(n, k, dtype=torch.float64, device="cuda")       # column-major k x n
    Cm = torch.rand(n, m, dtype=torch.float64, device="cuda")
    C0 = Cm.clone()
    st = torch.cuda.ExternalStream(p.stream(0))
    with torch.cuda.stream(st):
        p.execute_device(0, n, -0.7, B.data_ptr(), 0.8, Cm.data_ptr())
    torch.cuda.synchronize()
    # parity on a few columns against the oracle
    cols = sorted(set([0, 1, n // 2, n - 1]))
    Bh = B[cols].cpu().numpy().T.copy(order="F")
    Ch = C0[cols].cpu().numpy().T.copy(order="F")
    got = Cm[cols].cpu().numpy().T
    want = oracle.csrmm(rp32, c["col"], c["val"], Bh, -0.7, 0.8, Ch)
    bound = oracle.csrmm_bound(rp32, c["col"], c["val"], Bh, -0.7, 0.8, Ch)
    worst = float((np.abs(got - want) / np.maximum(bound, 1e-300)).max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        for _ in range(2):
            p.execute_device(0, n, -0.7, B.data_ptr(), 0.8, Cm.data_ptr())
        e0.record(st)
        for _ in range(reps):
            p.execute_device(0, n, -0.7, B.data_ptr(), 0.8, Cm.data_ptr())
        e1.record(st)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("spmm %s m=%d k=%d nnz=%d n=%d: %.3f ms  %.1f GFLOP/s  (plan %.2f s)  max err/bound %.2e" % (
        wl, m, k, nnz, n, ms, 2.0 * nnz * n / ms / 1e6, t_plan, worst))
    # cuSPARSE SpMM through torch for comparison (library baseline, diagnostic only)
    try:
        A_t = torch.sparse_csr_tensor(torch.from_numpy(c["rp"]), torch.from_numpy(c["col"].astype(np.int64)),
                                      torch.from_numpy(c["val"]), size=(m, k), dtype=torch.float64, device="cuda")
        Bd = B.T.contiguous()           # k x n row-major
        for _ in range(2):
            A_t @ Bd
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            A_t @ Bd
        torch.cuda.synchronize()
        ms2 = (time.perf_counter() - t0) / reps * 1e3
        print("cusparse (torch sparse_csr @ dense row-major): %.3f ms  %.1f GFLOP/s" % (ms2, 2.0 * nnz * n / ms2 / 1e6))
    except Exception as ex:
        print("cusparse comparison unavailable:", ex)
    p.destroy()


if __name__ == "__main__":
    main()
