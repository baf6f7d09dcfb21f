#!/usr/bin/env python3
"""run_test.py -- Python 3 port of what the reference driver actually runs (run_test.py:179-186 of
pnnl/s-blas calls test_spmv() and test_spmm(); the original is Python 2 and also carries disabled
sptrsv/sptrans halves, which are outside this repo's path).  For every matrix in matrices.txt and every
GPU count 1..n_gpus it runs

    ./test_spmv f <mtxpath><matrix> <gpu> 1 1 f          (run_test.py:45-66)
    ./test_spmm <mtxpath><matrix> 128 <gpu> 1            (run_test.py:160-175)

scrapes the `m:` / `Average` lines exactly like parse_spmv (run_test.py:29-43) and the `Matrix A --` /
`Matrix B --` / `SPMM:` lines like parse_spmm (run_test.py:146-159) and writes results.csv with the
reference's columns (V1/V2/V3 = baseline/v1/v2 for spmv).  Extra columns (gflops, alg_gbs) are appended
after the reference's so existing readers keep working.
"""
import os
import subprocess
import sys

n_gpus = int(os.environ.get("SBLAS_NGPUS", "2"))
mtxpath = os.environ.get("SBLAS_MTXPATH", "./sample_matrix/")
HERE = os.path.dirname(os.path.abspath(__file__))


def ensure_sample_matrix():
    """The reference bundles sample_matrix/qh768.mtx; here it is regenerated from the golden
    fixture (same entries, same order) when missing."""
    p = os.path.join(mtxpath, "qh768.mtx")
    if os.path.exists(p):
        return
    import numpy as np
    g = np.load(os.path.join(HERE, "tests", "golden", "qh768_coo.npz"))
    os.makedirs(mtxpath, exist_ok=True)
    with open(p, "w") as fh:
        fh.write("%%MatrixMarket matrix coordinate real general\n")
        fh.write("% Bai/qh768 regenerated from tests/golden/qh768_coo.npz (file order preserved)\n")
        fh.write("%d %d %d\n" % (int(g["m"]), int(g["n"]), len(g["val"])))
        for r, c, v in zip(g["row"], g["col"], g["val"]):
            fh.write("%d %d %s\n" % (r + 1, c + 1, repr(float(v))))


def parse_spmv(result):
    m = n = nnz = 0
    v1_time = v2_time = v3_time = float("nan")
    for line in result.strip().split("\n"):
        l = line.strip()
        if l.startswith("m:"):
            words = l.strip("\n").split(" ")
            m = int(words[1])
            n = int(words[3])
            nnz = int(words[5])
        if l.startswith("Average"):
            words = [i for i in l.strip("\n").split(" ") if i]
            v1_time = float(words[1])
            v2_time = float(words[2])
            v3_time = float(words[3])
    return m, n, nnz, v1_time, v2_time, v3_time


def test_spmv(mtxlist, result_file):
    result_file.write("kernel, matrix, n_gpu, m, n, nnz, version, time, gflops, alg_gbs\n")
    for mtx in mtxlist:
        for gpu in range(1, n_gpus + 1):
            cmd = "./test_spmv f " + mtxpath + mtx + " " + str(gpu) + " 1 " + "1 f"
            print(cmd)
            result = subprocess.run(cmd, shell=True, capture_output=True, text=True, cwd=HERE).stdout
            m, n, nnz, v1_time, v2_time, v3_time = parse_spmv(result)
            for label, t in (("V1", v1_time), ("V2", v2_time), ("V3", v3_time)):
                ok = t == t and t > 0
                alg = 12.0 * nnz + 4.0 * (m + 1) + 8.0 * n + 16.0 * m       # BASELINE.md section 2 (beta != 0)
                result_file.write("spmv, %s, %d, %d, %d, %d, %s, %s, %s, %s\n" % (
                    mtx, gpu, m, n, nnz, label, t, (2.0 * nnz / t / 1e9) if ok else "nan",
                    (alg / t / 1e9) if ok else "nan"))
    result_file.write("\n")


def parse_spmm(result):
    """run_test.py:146-159, token for token."""
    m = n = k = nnz = 0
    v1_time = float("nan")
    for line in result.strip().split("\n"):
        l = line.strip()
        if l.startswith("Matrix A --"):
            words = l.strip("\n").split(" ")
            m = int(words[4])
            n = int(words[6])
            nnz = int(words[8])
        if l.startswith("Matrix B --"):
            words = l.strip("\n").split(" ")
            k = int(words[6])
        if l.startswith("SPMM:"):
            words = l.strip("\n").split(" ")
            v1_time = float(words[5])
    return m, n, k, nnz, v1_time


def test_spmm(mtxlist, result_file):
    result_file.write("kernel, matrix, n_gpu, m, n, k, nnz, version, time, gflops\n")
    for mtx in mtxlist:
        for gpu in range(1, n_gpus + 1):
            cmd = "./test_spmm " + mtxpath + mtx + " 128 " + str(gpu) + " 1 "
            print(cmd)
            result = subprocess.run(cmd, shell=True, capture_output=True, text=True, cwd=HERE).stdout
            m, n, k, nnz, v1_time = parse_spmm(result)
            ok = v1_time == v1_time and v1_time > 0
            result_file.write("spmm, %s, %d, %d, %d, %d, %d, V1, %s, %s\n" % (
                mtx, gpu, m, n, k, nnz, v1_time, (2.0 * nnz * k / v1_time / 1e9) if ok else "nan"))
    result_file.write("\n")


def main():
    listing = os.path.join(HERE, "matrices.txt")
    mtxlist = [l.strip() for l in open(listing)] if os.path.exists(listing) else ["qh768.mtx"]
    mtxlist = [l for l in mtxlist if l]
    print("The following matrices will be tested:")
    print(mtxlist)
    ensure_sample_matrix()
    with open(os.path.join(HERE, "results.csv"), "w") as fh:
        test_spmv(mtxlist, fh)
        test_spmm(mtxlist, fh)


if __name__ == "__main__":
    sys.exit(main())
