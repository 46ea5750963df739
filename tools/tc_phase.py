import os, sys
os.environ["OAC_TC_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_gemm import run_gemm
for path in (1, 2):
    for (at, bt, M, N, K) in ((0, 0, 512, 256, 256), (0, 1, 256, 256, 8), (0, 1, 256, 256, 256), (1, 1, 256, 393, 256)):
        print("path", path, "layout", (at, bt), "M N K", M, N, K, flush=True)
        run_gemm(path, at, bt, M, N, K)
