"""Epilogue-cost probe: K=64 GEMMs (mainloop negligible) for the ViT output shapes and epilogue variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_features_b200  # noqa

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

M = 12000
tag = os.environ.get('VF_GEMM', 'pair')
for K in (64, 768):
    for N in (3072, 768):
        a = (torch.randn(M, K, device="cuda") * 0.1).half()
        b = (torch.randn(N, K, device="cuda") * 0.1).half()
        bias = torch.randn(N, device="cuda")
        res = torch.randn(M, N, device="cuda")
        t_plain16 = timeit(lambda: torch.ops.vfeat.gemm_f16(a, b, None, None, 0, False))
        t_bias16 = timeit(lambda: torch.ops.vfeat.gemm_f16(a, b, bias, None, 0, False))
        t_gelu16 = timeit(lambda: torch.ops.vfeat.gemm_f16(a, b, bias, None, 1, False))
        t_plain32 = timeit(lambda: torch.ops.vfeat.gemm_f16(a, b, None, None, 0, True))
        t_res32 = timeit(lambda: torch.ops.vfeat.gemm_f16(a, b, bias, None, 0, True))
        print(f"{tag:5s} K={K:4d} N={N:4d}: plain16 {t_plain16:6.1f} bias16 {t_bias16:6.1f} gelu16 {t_gelu16:6.1f} plain32 {t_plain32:6.1f} bias+res32 {t_res32:6.1f} us", flush=True)
