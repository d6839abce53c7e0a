"""Per-tile timeline of the tcgen05 GEMM from in-kernel SM-clock timestamps (debug build with -DVF_DBG_TRACE,
scripts/build_variants.sh -> libvfeat_trace.so).  For each ViT shape: when does the MMA thread get its accumulator stage,
when has it issued the tile's last MMA, when does the epilogue see the accumulator, when is it done -- i.e. who waits
for whom.   VF_LIBVFEAT=video_features_b200/libvfeat_trace.so python scripts/gemm_trace.py"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_features_b200 import _lib
l = _lib.lib()
raw = C.CDLL(_lib.LIB_PATH)
shapes = [(12500, 2304, 768, 0, 0, "QKV"), (12500, 768, 768, 0, 0, "out-proj"), (12500, 3072, 768, 0, 1, "fc1+GELU"),
          (12500, 768, 3072, 0, 0, "fc2"), (12500, 768, 768, 1, 0, "out-proj fp32")]
for (M, N, K, f32, act, name) in shapes:
    a = (torch.randn(M, K, device="cuda") * 0.1).half(); b = (torch.randn(N, K, device="cuda") * 0.1).half()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.float16)
    def run():
        rc = l.vf_gemm_f16(a.data_ptr(), K, b.data_ptr(), K, M, N, K, out.data_ptr(), N, f32, bias.data_ptr(), None, act,
                           torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    raw.vf_dbg_gemm_trace_clear()
    run()
    buf = np.zeros((74, 64, 8), dtype=np.int64)
    assert raw.vf_dbg_gemm_trace(buf.ctypes.data_as(C.c_void_p)) == 0
    tiles = ((M + 255) // 256) * ((N + 255) // 256)
    iters = (tiles + 73) // 74
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    entry, exit_ = buf[:, 63, 6].astype(np.float64), buf[:, 63, 7].astype(np.float64)
    g0, g1 = buf[:, 63, 4].astype(np.float64), buf[:, 63, 5].astype(np.float64)
    span_ns = g1.max() - g0.min()
    cyc = np.median(exit_ - entry)
    print(f"== {name}: {M}x{N}x{K}, {tiles} tiles, {iters} per pair; back-to-back launch period {us:.1f} us; first CTA entry -> "
          f"last CTA exit {span_ns / 1e3:.1f} us; median CTA life {cyc:.0f} cycles ({cyc / max(span_ns, 1) * 1e3:.0f} MHz)")
    print("   (cycles after the CTA's entry, median over the 74 pairs)")
    print("   iter | MMA gets stage | MMA issued all | epi sees acc | epi done | prod first..last load | period")
    prev = None
    full = buf[:, :, 1] > 0
    for it in range(min(iters, 10)):
        ok = full[:, it]
        if not ok.any():
            break
        t = buf[ok, it, :].astype(np.float64) - entry[ok, None]
        rel = np.median(t, axis=0)
        per = "" if prev is None else f"{rel[2] - prev:9.0f}"
        prev = rel[2]
        print(f"   {it:4d} | {rel[0]:12.0f} | {rel[1]:12.0f} | {rel[2]:12.0f} | {rel[3]:9.0f} | {rel[4]:9.0f} .. {rel[5]:9.0f} | {per}   ({int(ok.sum())} pairs)")
    last = np.array([buf[p, :63, 3].max() for p in range(74)], dtype=np.float64)
    print(f"   entry -> first load {np.median(buf[:, 0, 4] - entry):.0f}; last epilogue done -> exit {np.median(exit_ - last):.0f} cycles")
