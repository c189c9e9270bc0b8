"""Timing of the backward's GEMM dispatch (cneus_gemm_test) at training-step shapes: P = 131072 points."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from color_neus_b200 import _lib as L
lib = L.lib()
ws = torch.empty(lib.cneus_gemm_test_workspace_bytes(), dtype=torch.uint8, device="cuda")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
st = torch.cuda.current_stream().cuda_stream

def t(name, mode, M, N, K, lda, ldb, ldc, bias=False, relu=False, mask=False, acc=False, use_tc=1, reps=10):
    if mode == 2:
        A = torch.randn(K, lda, device="cuda") * 1e-5; B = torch.randn(K, ldb, device="cuda")
    else:
        A = torch.randn(M, lda, device="cuda"); B = torch.randn(N if mode == 0 else K, ldb, device="cuda")
    C = torch.zeros(M, ldc, device="cuda")
    bv = torch.randn(N, device="cuda") if bias else None
    mk = torch.randn(M, ldc, device="cuda") if mask else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    def run():
        L.check(lib.cneus_gemm_test(mode, A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, lda, ldb, ldc,
                                    bv.data_ptr() if bias else None, int(relu), mk.data_ptr() if mask else None, ldc, int(acc), use_tc,
                                    ws.data_ptr(), ws.numel(), st), "gemm")
    run(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    ms = tot / reps
    flop = 2.0 * M * N * K
    byt = 4.0 * ((M * K + M * N) if mode != 2 else (K * M + K * N))
    print(f"{name:44s} tc={use_tc} {ms*1e3:8.1f} us  {flop/ms/1e9:7.1f} TFLOP/s  {byt/ms/1e6:7.1f} GB/s (A+C or A+B)")

for tc in (1, 0):
    t("NT hidden 256x256 bias+relu", 0, P, 256, 256, 256, 256, 256, bias=True, relu=True, use_tc=tc)
    t("NT plain 256x256", 0, P, 256, 256, 256, 256, 256, use_tc=tc)
    t("NN 256x256", 1, P, 256, 256, 256, 256, 256, use_tc=tc)
    t("NN 256x256 mask", 1, P, 256, 256, 256, 256, 256, mask=True, use_tc=tc)
    t("TN 256x256", 2, 256, 256, P, 256, 256, 256, acc=True, use_tc=tc)
    t("NT first layer K=39", 0, P, 256, 39, 39, 39, 256, bias=True, use_tc=tc)
t("small NT N=3", 0, P, 3, 256, 256, 256, 3, bias=True)
t("small TN M=3", 2, 3, 256, P, 3, 256, 256, acc=True)
t("small-K NN K=3", 1, P, 256, 3, 3, 256, 256)
