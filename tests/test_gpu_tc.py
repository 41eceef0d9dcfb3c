"""tcgen05 / TMEM / bulk-TMA building blocks (bgx_tc_selftest) against exact integer products."""

import ctypes as C

import pytest
import torch

from bgflow_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(mode, A, W):
    lib = _lib.load()
    K = A.shape[1]
    scratch = torch.empty(128 * K, device=DEV)
    out = torch.full((128, 128), float("nan"), device=DEV)
    status = torch.zeros(4, dtype=torch.int32, device=DEV)
    rc = lib.bgx_tc_selftest(mode, A.data_ptr(), W.data_ptr(), K, scratch.data_ptr(), out.data_ptr(),
                             status.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "bgx_tc_selftest")
    torch.cuda.synchronize()
    assert int(status[0].item()) == 0, "an mbarrier wait timed out inside the kernel"
    return out


def _describe(out, ref):
    bad = (out != ref)
    rows = bad.any(dim=1).nonzero().flatten().tolist()
    cols = bad.any(dim=0).nonzero().flatten().tolist()
    return (f"{int(bad.sum())} of {bad.numel()} wrong; rows {rows[:8]}.. ({len(rows)}), cols {cols[:8]}.. "
            f"({len(cols)}); out[0,:4]={out[0, :4].tolist()} ref[0,:4]={ref[0, :4].tolist()}; "
            f"nan={int(torch.isnan(out).sum())}")


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("K", [32, 64, 128])
def test_exact_integer_gemm(mode, K):
    g = torch.Generator().manual_seed(K + mode)
    A = torch.randint(-3, 4, (128, K), generator=g).float().to(DEV)
    W = torch.randint(-3, 4, (128, K), generator=g).float().to(DEV)
    out = _run(mode, A, W)
    ref = A @ W.t()
    assert torch.equal(out, ref), f"mode {mode} K {K}: " + _describe(out, ref)


@pytest.mark.parametrize("mode", [0, 1])
def test_k_and_row_mapping(mode):
    """A = one-hot rows: out[m][n] must equal W[n][m % K] (catches any permutation of k / rows)."""
    K = 64
    A = torch.zeros(128, K, device=DEV)
    A[torch.arange(128), torch.arange(128) % K] = 1.0
    W = (torch.arange(128 * K, device=DEV).reshape(128, K) % 251).float()
    out = _run(mode, A, W)
    ref = W[:, torch.arange(128, device=DEV) % K].t().contiguous()
    assert torch.equal(out, ref), f"mode {mode}: " + _describe(out, ref)


@pytest.mark.parametrize("mode", [1])
def test_tf32_truncation_and_3x_split(mode):
    """Real-valued data: 1xTF32 is ~1e-3 accurate; hi/lo splitting of A and W recovers fp32."""
    g = torch.Generator().manual_seed(7)
    A = torch.randn(128, 128, generator=g).to(DEV)
    W = (torch.randn(128, 128, generator=g) * 0.1).to(DEV)
    ref = (A.double() @ W.double().t())
    out1 = _run(mode, A, W)
    err1 = (out1.double() - ref).abs().max().item()
    assert err1 < 2e-2, err1

    def split(x):
        hi = (x.view(torch.int32) & -8192).view(torch.float32)
        return hi, x - hi
    a_hi, a_lo = split(A)
    w_hi, w_lo = split(W)
    out3 = _run(mode, a_hi, w_hi) + _run(mode, a_lo, w_hi) + _run(mode, a_hi, w_lo)
    err3 = (out3.double() - ref).abs().max().item()
    assert err3 < 2e-5, (err1, err3)


def test_tmem_load_at_unaligned_column_offset():
    """mode 2: tcgen05.ld 32x32b.x32 starting at columns 25 and 75 (not multiples of 32)."""
    g = torch.Generator().manual_seed(11)
    A = torch.randint(-3, 4, (128, 64), generator=g).float().to(DEV)
    W = torch.randint(-3, 4, (128, 64), generator=g).float().to(DEV)
    out = _run(2, A, W)
    ref = A @ W.t()
    assert torch.equal(out[:, :32], ref[:, 25:57]), "offset 25"
    assert torch.equal(out[:, 32:64], ref[:, 75:107]), "offset 75"


@pytest.mark.parametrize("K", [64, 128])
def test_bf16_ts_exact_integer_gemm(K):
    """mode 3: the production operand path (bf16 pairs in TMEM x bf16 tiles in smem, fp32 accumulate)."""
    g = torch.Generator().manual_seed(100 + K)
    A = torch.randint(-3, 4, (128, K), generator=g).float().to(DEV)
    W = torch.randint(-3, 4, (128, K), generator=g).float().to(DEV)
    out = _run(3, A, W)
    ref = A @ W.t()
    assert torch.equal(out, ref), f"bf16 TS K {K}: " + _describe(out, ref)


def test_bf16_ts_k_and_row_mapping():
    K = 64
    A = torch.zeros(128, K, device=DEV)
    A[torch.arange(128), torch.arange(128) % K] = 1.0
    W = (torch.arange(128 * K, device=DEV).reshape(128, K) % 251).float()   # < 256: exact in bf16
    out = _run(3, A, W)
    ref = W[:, torch.arange(128, device=DEV) % K].t().contiguous()
    assert torch.equal(out, ref), "bf16 TS: " + _describe(out, ref)
