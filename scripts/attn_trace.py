"""Per-tile timeline of the fused QKV + attention kernel (debug build with -DVF_DBG_TRACE)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_features_b200 import _lib, synthetic_weights
from video_features_b200.clip_engine import ClipEngine
raw = C.CDLL(_lib.LIB_PATH)
eng = ClipEngine(synthetic_weights.clip_vit_b32_state_dict(0), device=0)
x = torch.randn(250 * 50, 768, device="cuda").half()
for _ in range(3):
    eng.block_attention(3, x, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    eng.block_attention(3, x, True)
e1.record(); torch.cuda.synchronize()
print(f"fused QKV+attention, 250 frames: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch (600 tiles, 8.1 per pair)")
e0.record()
for _ in range(20):
    eng.block_attention(3, x, False)
e1.record(); torch.cuda.synchronize()
print(f"split (QKV GEMM + attention kernel): {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
eng.block_attention(3, x, True)
buf = np.zeros((74, 64, 8), dtype=np.int64)
assert raw.vf_dbg_attn_trace(buf.ctypes.data_as(C.c_void_p)) == 0
print("iter | tile staged | unit done | stored   (cycles after tile 0 was staged, median over pairs; attention warp 8 of the leader CTA)")
t0 = buf[:, 0, 3].astype(np.float64)[:, None]
for it in range(9):
    rel = np.median(buf[:, it, 3:6].astype(np.float64) - t0, axis=0)
    print(f"{it:4d} | " + " | ".join(f"{v:9.0f}" for v in rel))
d = buf[:, 1:7, :6].astype(np.float64)
print(f"   attention unit: {np.median(d[:, :, 4] - d[:, :, 3]):.0f} cycles; store: {np.median(d[:, :, 5] - d[:, :, 4]):.0f}; "
      f"wait for the next staged tile: {np.median(d[:, 1:, 3] - d[:, :-1, 5]):.0f}; tile period: {np.median(d[:, 1:, 3] - d[:, :-1, 3]):.0f} cycles")
print("loader warp 4:  iter | acc ready | buffer free | staged")
for it in range(8):
    rel = np.median(buf[:, it, 0:3].astype(np.float64) - t0, axis=0)
    print(f"               {it:4d} | " + " | ".join(f"{v:9.0f}" for v in rel))
print(f"   loader: wait buffer {np.median(d[:, :, 1] - d[:, :, 0]):.0f}, stage {np.median(d[:, :, 2] - d[:, :, 1]):.0f} cycles; acc ready period {np.median(d[:, 1:, 0] - d[:, :-1, 0]):.0f}")
