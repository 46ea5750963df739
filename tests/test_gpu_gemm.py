"""The two GEMM stage kernels in isolation (fp32 SIMT and tcgen05 kind::tf32) against torch fp32/fp64:
every operand-layout combination the step uses (forward X W^T, dX = dY W, dW = dY^T X), ragged edges."""
import ctypes as C

import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def run_gemm(path, a_trans, b_trans, M, N, K, bias=False, relu=False, pad=0, seed=0, align4=False, want_kernel=None):
    from oac_explore_b200 import _lib
    g = torch.Generator(device='cuda').manual_seed(seed)
    lda = (M if a_trans else K) + pad
    ldb = (N if b_trans else K) + pad
    ldc = N + pad
    if align4:          # 16-byte row strides: what the TMA-fed kernel needs (the step's own buffers are laid out so)
        lda, ldb, ldc = [(x + 3) // 4 * 4 for x in (lda, ldb, ldc)]
    A = torch.randn((K if a_trans else M, lda), device='cuda', generator=g)
    B = torch.randn((K if b_trans else N, ldb), device='cuda', generator=g)
    Cm = torch.full((M, ldc), 7.0, device='cuda')
    bvec = torch.randn(N, device='cuda', generator=g) if bias else None
    _lib.check(_lib.lib().oac_gemm_debug(path, int(a_trans), int(b_trans), M, N, K, _lib.ptr(A), lda, _lib.ptr(B), ldb,
                                         _lib.ptr(Cm), ldc, _lib.ptr(bvec), int(relu), _lib.current_stream()),
               "oac_gemm_debug")
    if want_kernel is not None:
        assert _lib.lib().oac_gemm_debug_kernel() == want_kernel
    Am = (A[:, :M].t() if a_trans else A[:, :K]).double()
    Bm = (B[:, :N].t() if b_trans else B[:, :K]).double()
    ref = Am @ Bm.t()
    if bias:
        ref = ref + bvec.double()
    if relu:
        ref = torch.relu(ref)
    assert torch.all(Cm[:, N:] == 7.0), "wrote outside the N columns"
    return Cm[:, :N].double().cpu(), ref.cpu()


SHAPES = [(256, 256, 256), (512, 256, 376), (256, 256, 393), (256, 34, 256), (256, 1, 256), (256, 17, 256),
          (256, 393, 256), (34, 256, 256), (1, 256, 256), (100, 70, 50), (128, 16, 8), (130, 33, 133)]
LAYOUTS = [(0, 0), (0, 1), (1, 1)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("a_trans,b_trans", LAYOUTS)
def test_simt_gemm(M, N, K, a_trans, b_trans):
    for pad in (0, 3):
        got, ref = run_gemm(0, a_trans, b_trans, M, N, K, bias=(a_trans == 0), relu=(b_trans == 0 and a_trans == 0), pad=pad)
        assert rel_err(got, ref) <= 2e-6, (pad, rel_err(got, ref))


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("a_trans,b_trans", LAYOUTS)
def test_tcgen05_tf32_gemm(M, N, K, a_trans, b_trans):
    for pad in (0, 3):
        got, ref = run_gemm(1, a_trans, b_trans, M, N, K, bias=(a_trans == 0), relu=False, pad=pad)
        # TF32 operands (10-bit mantissa), fp32 accumulate: ~5e-4 relative on random data
        assert rel_err(got, ref) <= 1.5e-3, (pad, rel_err(got, ref))


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("a_trans,b_trans", LAYOUTS)
def test_tcgen05_3xtf32_gemm(M, N, K, a_trans, b_trans):
    for pad in (0, 3):
        got, ref = run_gemm(2, a_trans, b_trans, M, N, K, bias=(a_trans == 0), relu=False, pad=pad)
        assert rel_err(got, ref) <= 3e-6, (pad, rel_err(got, ref))


WS_SHAPES = SHAPES + [(1024, 256, 393), (384, 512, 264), (256, 393, 1000), (640, 224, 96)]


@pytest.mark.parametrize("M,N,K", [(256, 256, 256), (512, 256, 376), (256, 256, 393), (256, 34, 256), (256, 1, 256),
                                   (256, 393, 256), (1024, 256, 393), (256, 393, 1000), (768, 192, 40)])
@pytest.mark.parametrize("a_trans,b_trans", LAYOUTS)
def test_ws2_cta_pair_gemm(M, N, K, a_trans, b_trans, monkeypatch):
    """CTA-pair kernel (gemm_ws2.cuh: tcgen05.mma.cta_group::2, 256-row tiles, each CTA loads half of the B tile) on every
    operand layout: OAC_WS2_ALL=1 lifts the planner's restriction to the stages where pairs pay, so shapes made of whole
    256-row tiles run on it, including ragged N / K, N = 1 and more K chunks than ring slots."""
    monkeypatch.setenv("OAC_WS2_ALL", "1")
    for relu in (False, True):
        got, ref = run_gemm(1, a_trans, b_trans, M, N, K, bias=(a_trans == 0), relu=relu and a_trans == 0, align4=True,
                            want_kernel=3)
        assert rel_err(got, ref) <= 1.5e-3, rel_err(got, ref)


@pytest.mark.parametrize("M,N,K", WS_SHAPES)
@pytest.mark.parametrize("a_trans,b_trans", LAYOUTS)
def test_ws_tcgen05_tf32_gemm(M, N, K, a_trans, b_trans):
    """Warp-specialised TMA + tcgen05 kernel (operands rounded to tf32 by the TMA unit), all operand layouts, ragged
    M / N / K edges, multi-tile M and N, more K chunks than ring slots."""
    for relu in (False, True):
        got, ref = run_gemm(1, a_trans, b_trans, M, N, K, bias=(a_trans == 0), relu=relu and a_trans == 0, align4=True,
                            want_kernel=2)
        assert rel_err(got, ref) <= 1.5e-3, rel_err(got, ref)


@pytest.mark.parametrize("M,N,K", WS_SHAPES)
@pytest.mark.parametrize("a_trans,b_trans", LAYOUTS)
def test_simt_gemm_tma_staged_operands(M, N, K, a_trans, b_trans):
    """16-byte aligned operands make the FFMA tile fetch its whole-K tiles with TMA (fp32 tensor maps, SWIZZLE_128B atoms
    for K-contiguous operands, bounds zero-filled by the TMA unit); K = 1000 exceeds the shared-memory budget and takes
    the chunked cp.async path through the same kernel."""
    got, ref = run_gemm(0, a_trans, b_trans, M, N, K, bias=(a_trans == 0), relu=(b_trans == 0 and a_trans == 0), align4=True,
                        want_kernel=0)
    assert rel_err(got, ref) <= 2e-6, rel_err(got, ref)
