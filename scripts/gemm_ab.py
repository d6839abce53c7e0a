"""A/B timing of vf_gemm_f16 from two builds of the library on the CLIP tower's GEMM shapes (development aid).
usage: gemm_ab.py libA.so libB.so"""
import ctypes as C, sys, torch
torch.cuda.init()
shapes = [  # M, N, K, out_f32   (a 250-frame chunk: QKV, out-proj, fc1, fc2)
    (12500, 2304, 768, 0), (12500, 768, 768, 0), (12500, 3072, 768, 0), (12500, 768, 3072, 0), (12500, 768, 768, 1)]
libs = []
for p in sys.argv[1:]:
    l = C.CDLL(p)
    l.vf_gemm_f16.restype = C.c_int
    l.vf_gemm_f16.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    libs.append((p, l))
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rep in range(2):
    for (M, N, K, f32) in shapes:
        a = (torch.randn(M, K, device="cuda") * 0.1).half(); b = (torch.randn(N, K, device="cuda") * 0.1).half()
        bias = torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.float16)
        row = []
        for p, l in libs:
            def run():
                rc = l.vf_gemm_f16(a.data_ptr(), K, b.data_ptr(), K, M, N, K, out.data_ptr(), N, f32, bias.data_ptr(), None, 1 if N == 3072 else 0,
                                   torch.cuda.current_stream().cuda_stream)
                assert rc == 0
            ms = timeit(run)
            row.append(f"{p.split('/')[-1]}: {ms*1e3:7.1f} us {2.0*M*N*K/ms/1e9:7.1f} TF")
        print(f"{M}x{N}x{K} f32={f32} | " + " | ".join(row), flush=True)
