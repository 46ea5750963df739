"""Two gemm_ws launches for an ncu capture: layouts (0,0) and (0,1) at M = 4 x 148 x 128, N = 256, K = 64 (epilogue-dominated)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oac_explore_b200 import _lib
L = _lib.lib()
M, N, K = 4 * 148 * 128, 256, int(os.environ.get("K", "64"))
for bt in (0, 1):
    A = torch.randn((M, K), device='cuda'); B = torch.randn((K if bt else N, N if bt else K), device='cuda')
    C = torch.empty((M, N), device='cuda')
    _lib.check(L.oac_gemm_debug(1, 0, bt, M, N, K, _lib.ptr(A), K, _lib.ptr(B), N if bt else K, _lib.ptr(C), N, None, 0,
                                _lib.current_stream()), "gemm")
    torch.cuda.synchronize()
