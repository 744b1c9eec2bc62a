"""Library comparison line (not the product path): cuBLAS bf16 throughput on the encoder's GEMM shapes."""
import torch
M = 18944
for (N, K) in [(1152, 384), (384, 384), (1536, 384), (384, 1536), (2304, 768), (768, 768), (3072, 768), (768, 3072)]:
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    for _ in range(5):
        torch.matmul(a, w.t())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        torch.matmul(a, w.t())
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(f"M={M} N={N} K={K}: {us:.1f} us  {2*M*N*K/us/1e6:.0f} TFLOP/s")
