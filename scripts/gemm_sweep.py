"""Mainloop-rate probe: large square fp16 GEMMs through vf_gemm_f16 vs torch.matmul (cuBLAS) on the same shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_features_b200  # noqa

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

shapes = [(8192, 8192, 8192), (12000, 768, 3072), (12000, 3072, 768), (12000, 2304, 768), (12000, 768, 768)]
for (M, N, K) in shapes:
    a = (torch.randn(M, K, device="cuda") * 0.1).half()
    b = (torch.randn(N, K, device="cuda") * 0.1).half()
    ms = timeit(lambda: torch.ops.vfeat.gemm_f16(a, b, None, None, 0, False))
    ms_t = timeit(lambda: a @ b.t())
    fl = 2.0 * M * N * K
    print(f"{os.environ.get('VF_GEMM','pair'):5s} {M}x{N}x{K}: vf {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF | cuBLAS {ms_t*1e3:8.1f} us {fl/ms_t/1e9:7.1f} TF", flush=True)
