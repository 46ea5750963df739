"""One round of 128 x 256 tiles on 8 .. 148 CTAs: does a tile get faster when fewer SMs pull operands at the same time?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("OAC_GEMM_DEBUG_REPS", "20")
import torch
from oac_explore_b200 import _lib
L = _lib.lib()
def run(M, N, K):
    A = torch.randn((M, K), device='cuda'); B = torch.randn((N, K), device='cuda'); C = torch.empty((M, N), device='cuda')
    bv = torch.randn(N, device='cuda')
    _lib.check(L.oac_gemm_debug(1, 0, 0, M, N, K, _lib.ptr(A), K, _lib.ptr(B), K, _lib.ptr(C), N, _lib.ptr(bv), 1, _lib.current_stream()), "gemm")
    torch.cuda.synchronize()
for M in (1024, 18944):
    for K in (256, 1024):
        run(M, 256, K)
