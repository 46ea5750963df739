"""Time ONE gemm_ws stage per operand layout at the shapes of the batched-seed step (measurement aid: the stage kernel is
driven through oac_gemm_debug with OAC_GEMM_DEBUG_REPS, which prints the mean launch time to stderr)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("OAC_GEMM_DEBUG_REPS", "20")
import torch
from oac_explore_b200 import _lib

L = _lib.lib()
def run(at, bt, M, N, K, bias=False, padb=0, pada=0):
    lda = (M if at else K) + pada
    ldb = (N if bt else K) + padb
    A = torch.randn((K if at else M, lda), device='cuda')
    B = torch.randn((K if bt else N, ldb), device='cuda')
    C = torch.empty((M, N), device='cuda')
    bv = torch.randn(N, device='cuda') if bias else None
    _lib.check(L.oac_gemm_debug(1, at, bt, M, N, K, _lib.ptr(A), lda, _lib.ptr(B), ldb, _lib.ptr(C), N, _lib.ptr(bv), int(bias),
                                _lib.current_stream()), "gemm")
    torch.cuda.synchronize()

import sys
M = 4 * 148 * 128
print("ldb pads on (0,1) N=256 K=256:", file=sys.stderr)
for padb in (0, 4, 32, 64, 132, 256):
    run(0, 1, M, 256, 256, padb=padb)
print("lda pads on (0,0) / (0,1):", file=sys.stderr)
for pada in (0, 4, 32, 140):
    run(0, 0, M, 256, 256, bias=True, pada=pada)
    run(0, 1, M, 256, 256, pada=pada)
print("K sweep (0,0) vs (0,1), N=256:", file=sys.stderr)
for K in (64, 128, 512):
    run(0, 0, M, 256, K, bias=True)
    run(0, 1, M, 256, K)
print("layouts at N=256 K=256, 1 and 4 rounds:", file=sys.stderr)
for M in (148 * 128, 4 * 148 * 128):
    run(0, 0, M, 256, 256, bias=True); run(0, 1, M, 256, 256); run(1, 1, M, 256, 256)
