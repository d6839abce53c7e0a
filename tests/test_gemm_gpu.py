"""tcgen05 GEMM (vf_gemm_f16 through the C ABI) against a plain fp32 torch reference of the same op."""
import pytest
import torch

from conftest import rel_l2

import video_features_b200  # noqa: F401  (registers torch.ops.vfeat)

pytestmark = pytest.mark.gpu


def _ref(a, b, bias, scale, residual, act):
    y = a.float() @ b.float().t()
    if scale is not None:
        y = y * scale
    if bias is not None:
        y = y + bias
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    elif act == 2:
        y = torch.relu(y)
    if residual is not None:
        y = y + residual
    return y


CASES = [
    # M, N, K, bias, scale, residual, act, out_f32
    (128, 256, 64, False, False, False, 0, True),
    (128, 256, 128, False, False, False, 0, True),
    (300, 256, 768, True, False, False, 0, True),
    (300, 128, 200, True, False, False, 0, True),      # K tail (zero-filled by TMA), BN=128
    (77, 64, 64, True, True, False, 2, False),         # BN=64, M tail, scale+relu, fp16 out
    (500, 96, 320, True, False, True, 0, True),        # N not a multiple of the tile
    (6000, 2304, 768, True, False, False, 0, False),   # ViT QKV
    (6000, 768, 768, True, False, True, 0, True),      # ViT out-proj + residual
    (6000, 3072, 768, True, False, False, 1, False),   # ViT fc1 + QuickGELU
    (6000, 768, 3072, True, False, True, 0, True),     # ViT fc2 + residual
    (120, 512, 768, False, False, False, 0, True),     # final projection
    (20000, 768, 768, True, False, True, 0, True),     # many tiles per CTA (persistent loop, phase wrap)
]


@pytest.mark.parametrize("M,N,K,has_bias,has_scale,has_res,act,out_f32", CASES)
def test_gemm_matches_fp32_reference(cuda_device, M, N, K, has_bias, has_scale, has_res, act, out_f32):
    from video_features_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, generator=g) * 0.5).half().to(cuda_device)
    b = (torch.randn(N, K, generator=g) * K ** -0.5).half().to(cuda_device)
    bias = torch.randn(N, generator=g).to(cuda_device) if has_bias else None
    scale = (1 + 0.1 * torch.randn(N, generator=g)).to(cuda_device) if has_scale else None
    res = torch.randn(M, N, generator=g).to(cuda_device) if has_res else None
    out = torch.ops.vfeat.gemm_f16(a, b, bias, scale, res, act, out_f32)
    torch.cuda.synchronize()
    ref = _ref(a, b, bias, scale, res, act)
    assert out.shape == (M, N)
    assert torch.isfinite(out.float()).all()
    tol = 2e-5 if out_f32 else 1.5e-3     # fp32 accumulation-order noise / fp16 output rounding
    err = rel_l2(out.float(), ref)
    assert err < tol, f"rel L2 {err:.3e}"
    mx = float((out.float() - ref).abs().max() / ref.abs().max())
    assert mx < (1e-4 if out_f32 else 2e-3), f"max err {mx:.3e}"


def test_gemm_rejects_bad_arguments(cuda_device):
    from video_features_b200._lib import VfError
    a = torch.zeros(16, 60, dtype=torch.float16, device=cuda_device)   # K not a multiple of 8
    b = torch.zeros(64, 60, dtype=torch.float16, device=cuda_device)
    with pytest.raises(VfError):
        torch.ops.vfeat.gemm_f16(a, b, None, None, None, 0, True)
