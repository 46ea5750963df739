import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_gemm import run_gemm
from tests.util import rel_err
for path in (0, 1, 2):
    for (at, bt) in ((0, 0), (0, 1), (1, 1)):
        row = []
        for K in (64, 128, 192, 256, 320, 384, 393, 512, 1024):
            got, ref = run_gemm(path, at, bt, 256, 64, K)
            row.append("%d:%.1e" % (K, rel_err(got, ref)))
        print("path", path, "layout", (at, bt), " ".join(row))
