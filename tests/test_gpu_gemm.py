"""GEMMs of the training backward (color_neus_b200/csrc/gemm.cu fp32 SGEMM, gemm_tc.cu tcgen05 split-precision) against
torch fp64 matmul: the three operand modes with the leading dimensions, column offsets, narrow shapes and epilogues
(bias, ReLU, mask, accumulate) backward.cu uses, including adjoint-sized (tiny) magnitudes.
Tolerance: max |err| <= 2e-5 * max |ref| (fp16 hi/lo 3-pass keeps ~22 bits; the bar for gradients is 5e-3)."""
import numpy as np
import pytest
import torch

from color_neus_b200 import _lib as L

pytestmark = pytest.mark.gpu


def run_gemm(mode, A, B, Cshape, M, N, K, lda, ldb, ldc, bias=None, relu=False, mask=None, ldmask=0, acc_init=None, use_tc=True):
    lib = L.lib()
    ws = torch.empty(lib.cneus_gemm_test_workspace_bytes(), dtype=torch.uint8, device="cuda")
    C = torch.zeros(Cshape, device="cuda") if acc_init is None else acc_init.clone()
    L.check(lib.cneus_gemm_test(mode, A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, lda, ldb, ldc,
                                bias.data_ptr() if bias is not None else None, int(relu),
                                mask.data_ptr() if mask is not None else None, ldmask, int(acc_init is not None), int(use_tc),
                                ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "cneus_gemm_test")
    torch.cuda.synchronize()
    return C


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300))


@pytest.mark.parametrize("use_tc", [True, False])
@pytest.mark.parametrize("M,N,K,lda,ldc,scale", [
    (4096, 256, 256, 256, 256, 1.0),      # hidden layer
    (2048 + 77, 217, 256, 256, 256, 1.0),  # skip layer, ragged M, padded ldc
    (4096, 256, 39, 39, 256, 1.0),        # first layer (unaligned rows of A)
    (4096, 257, 256, 256, 257, 1.0),      # last SDF layer (N > 256 -> split)
    (4096, 3, 256, 256, 3, 1.0),          # colour head (narrow)
    (4096, 256, 262, 262, 256, 1e-7),     # colour first layer, adjoint-sized magnitudes
    (4096, 256, 33, 33, 256, 1.0),        # relight input layer
])
def test_nt_matches_fp64(use_tc, M, N, K, lda, ldc, scale):
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, lda, generator=g) * scale * torch.exp(3 * torch.randn(M, 1, generator=g))).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda() * scale
    C = run_gemm(0, A, W, (M, ldc), M, N, K, lda, K, ldc, bias=bias, relu=True, use_tc=use_tc)
    ref = torch.relu(A[:, :K].double() @ W.double().T + bias.double())
    assert rel(C[:, :N], ref) < 2e-5


@pytest.mark.parametrize("use_tc", [True, False])
@pytest.mark.parametrize("M,N,K,ldb,boff,scale", [
    (4096, 256, 256, 256, 0, 1e-6),       # dL/dx of a hidden layer
    (4096, 256, 217, 256, 0, 1e-3),
    (4096, 39, 256, 39, 0, 1.0),          # first layer
    (4096, 262, 256, 262, 0, 1e-5),       # colour first layer (N > 256)
    (4096, 256, 256, 259, 3, 1e-4),       # relight re-injection layer: weight columns 3..258
    (4096, 3, 256, 259, 0, 1e-4),         # ... and its 3 colour columns (narrow), accumulated
    (4096, 256, 3, 256, 0, 1e-4),         # from a 3-wide adjoint (K = 3: SGEMM path)
])
def test_nn_matches_fp64(use_tc, M, N, K, ldb, boff, scale):
    g = torch.Generator().manual_seed(M + N + K + 1)
    A = (torch.randn(M, K, generator=g) * scale * torch.exp(3 * torch.randn(M, 1, generator=g))).cuda()
    W = (torch.randn(K, ldb, generator=g) / K ** 0.5).cuda()
    mask = torch.randn(M, N, generator=g).cuda()
    init = torch.randn(M, N, generator=g).cuda() * scale
    use_mask = N >= 16
    Wv = W[:, boff:]
    C = run_gemm(1, A, Wv, (M, N), M, N, K, K, ldb, N, mask=mask if use_mask else None, ldmask=N, acc_init=init, use_tc=use_tc)
    ref = A.double() @ W[:, boff:boff + N].double() + init.double()  # epilogue order: accumulate, (ReLU,) mask
    if use_mask:
        ref = torch.where(mask.double() > 0, ref, torch.zeros_like(ref))
    assert rel(C, ref) < 2e-5


@pytest.mark.parametrize("use_tc", [True, False])
@pytest.mark.parametrize("K,M,N,lda,ldb,scale", [
    (8192, 256, 256, 256, 256, 1e-6),
    (8192 + 100, 217, 256, 217, 256, 1e-6),
    (8192, 256, 39, 256, 39, 1e-6),
    (8192, 257, 256, 257, 256, 1e-6),     # last SDF layer: M > 256 -> 1 + 256
    (8192, 256, 262, 256, 262, 1e-6),     # N > 256 -> 6 + 256
    (8192, 3, 256, 3, 256, 1e-6),         # 3-wide adjoint (narrow)
    (8192, 256, 33, 256, 33, 1e-6),
    (150000, 256, 256, 256, 256, 1e-6),   # many splits
])
def test_tn_matches_fp64(use_tc, K, M, N, lda, ldb, scale):
    g = torch.Generator().manual_seed(M + N + K + 2)
    A = (torch.randn(K, lda, generator=g) * scale * torch.exp(3 * torch.randn(K, 1, generator=g))).cuda()
    B = torch.randn(K, ldb, generator=g).cuda()
    init = torch.randn(M, N, generator=g).cuda() * scale
    C = run_gemm(2, A, B, (M, N), M, N, K, lda, ldb, N, acc_init=init, use_tc=use_tc)
    ref = A[:, :M].double().T @ B[:, :N].double() + init.double()
    assert rel(C, ref) < 2e-5
