"""ncu target: epilogue-dominated GEMMs (K = 64) with and without the QuickGELU epilogue (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_features_b200  # noqa
M, N, K = 12000, 3072, int(sys.argv[1]) if len(sys.argv) > 1 else 64
a = (torch.randn(M, K, device="cuda") * 0.1).half(); b = (torch.randn(N, K, device="cuda") * 0.1).half()
bias = torch.randn(N, device="cuda")
for act in (0, 1, 0, 1):
    torch.ops.vfeat.gemm_f16(a, b, bias, None, act, False)
torch.cuda.synchronize()
