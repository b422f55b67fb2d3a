"""GPU: the tcgen05/TMEM conventions (descriptor bits, canonical no-swizzle staging, arbitrary row starts,
MN-major operands, TMEM lane mapping) on a single tile against a torch fp32 matmul of the same bf16 data."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(K, N, mode, Ra=128, row_off=0, grp_rows=8, seed=0):
    from musicgan_b200 import _lib, build
    l = ctypes.CDLL(build.DEBUG_LIB)          # test-only probes: their own library, not the product .so
    l.mg_debug_umma_gemm.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 6 + [ctypes.c_void_p]
    g = torch.Generator().manual_seed(seed)
    if mode == 0:
        A = torch.randn(Ra, K, generator=g).bfloat16().cuda()
        B = torch.randn(N, K, generator=g).bfloat16().cuda()
        rows = torch.tensor([row_off + (m // 8) * grp_rows + m % 8 for m in range(128)])
        ref = A.float()[rows.cuda()] @ B.float().t()
    elif mode == 1:
        A = torch.randn(K, 128, generator=g).bfloat16().cuda()
        B = torch.randn(K, N, generator=g).bfloat16().cuda()
        ref = A.float().t() @ B.float()
    else:
        A = torch.randn(128, K, generator=g).bfloat16().cuda()
        B = torch.randn(K, N, generator=g).bfloat16().cuda()
        ref = A.float() @ B.float()
    D = torch.zeros(128, N, device="cuda")
    _lib.check(l.mg_debug_umma_gemm(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, mode, Ra, row_off, grp_rows,
                                    torch.cuda.current_stream().cuda_stream), "mg_debug_umma_gemm")
    torch.cuda.synchronize()
    err = (D - ref).abs().max().item()
    scale = ref.abs().max().item()
    return err, scale


@pytest.mark.parametrize("K,N", [(16, 16), (32, 32), (64, 48), (160, 160), (288, 16)])
def test_k_major_dense(K, N):
    err, scale = _run(K, N, 0)
    assert err <= 2e-3 * scale, (err, scale)


@pytest.mark.parametrize("row_off,grp_rows", [(0, 10), (1, 10), (11, 10), (22, 10), (3, 8)])
def test_k_major_shifted_rows(row_off, grp_rows):
    """A convolution tap = the same staged tile read from another start row with a halo pitch of 10."""
    err, scale = _run(32, 32, 0, Ra=200, row_off=row_off, grp_rows=grp_rows)
    assert err <= 2e-3 * scale, (err, scale)


@pytest.mark.parametrize("K,N", [(16, 16), (128, 32), (256, 160)])
def test_mn_major(K, N):
    err, scale = _run(K, N, 1)
    assert err <= 2e-3 * scale, (err, scale)


@pytest.mark.parametrize("K,N", [(16, 16), (128, 32), (128, 144)])
def test_a_operand_from_tmem(K, N):
    """A rows in TMEM (lane = row, two bf16 per 32-bit column, low half = even k), B MN-major in smem."""
    err, scale = _run(K, N, 2)
    err_swapped, _ = _run(K, N, 3)
    print(f"TS mode K={K} N={N}: err {err:.3e} (swapped halves {err_swapped:.3e}), scale {scale:.3e}")
    assert err <= 2e-3 * scale, (err, err_swapped, scale)
