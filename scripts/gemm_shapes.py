"""Timing of vf_gemm_f16 on the CLIP tower's GEMM shapes (250-frame chunk), raw ctypes (development aid)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_features_b200 import _lib
l = _lib.lib()
shapes = [(12500, 2304, 768, 0, 0), (12500, 768, 768, 0, 0), (12500, 3072, 768, 0, 1), (12500, 768, 3072, 0, 0), (12500, 768, 768, 1, 0)]
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K, f32, act) in shapes:
    a = (torch.randn(M, K, device="cuda") * 0.1).half(); b = (torch.randn(N, K, device="cuda") * 0.1).half()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.float16)
    def run():
        rc = l.vf_gemm_f16(a.data_ptr(), K, b.data_ptr(), K, M, N, K, out.data_ptr(), N, f32, bias.data_ptr(), None, act,
                           torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    ms = timeit(run)
    print(f"{os.environ.get('VF_GEMM_EPI3', '-')} {M}x{N}x{K} f32={f32} act={act}: {ms*1e3:7.1f} us {2.0*M*N*K/ms/1e9:7.1f} TF", flush=True)
