"""The tcgen05 3xTF32 GEMM kernel against fp64 torch.matmul: fp32-faithful (1e-5 would be the
interaction-block budget; the kernel itself is held to 2e-6 relative to the row scale)."""
import pytest
import torch

from e3b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(out, ref):
    return float((out.double().cpu() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("M,N,K", [(1000, 1920, 64), (300, 64, 1920), (129, 200, 8), (128, 128, 32), (5000, 320, 192),
                                   (77, 64, 384), (1, 8, 4)])
def test_gemm_plain(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    ref = 0.37 * (A.double() @ B.double().T)
    C = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm_tf32x3(A.to(DEV), B.to(DEV), C, M, N, K, alpha=0.37)
    assert _rel(C, ref) < 2e-6


def test_gemm_affine_rows_and_strided_columns():
    """A rows (z, m) inside [z][D] node rows (imu block), C written in mul_ir layout [z][w][m]"""
    g = torch.Generator().manual_seed(1)
    Z, d, K, N, D = 301, 5, 64, 64, 1152
    off = 512                                   # block offset inside the node row
    X = torch.randn(Z, D, generator=g)
    B = torch.randn(N, K, generator=g)
    a = X[:, off:off + d * K].reshape(Z * d, K)
    ref = (a.double() @ B.double().T).reshape(Z, d, N).transpose(1, 2).reshape(Z, N * d)     # [z][w][m]
    Xd = X.to(DEV)
    C = torch.zeros(Z, N * d, device=DEV)
    ops.gemm_tf32x3(Xd[:, off:], B.to(DEV), C, Z * d, N, K, a_rows=(D, K, d), c_rows=(N * d, 1, d), c_col_stride=d)
    assert _rel(C, ref) < 2e-6


@pytest.mark.parametrize("V", [16, 32])
def test_gemm_reduce_epilogue(V):
    """self-connection: out[(z,m), w] = sum_{u,v} x[(z,m), u] W[u, v, w] a[z, v]"""
    g = torch.Generator().manual_seed(V)
    Z, d, U, Wn = 200, 3, 64, 40
    x = torch.randn(Z * d, U, generator=g)
    W = torch.randn(U, V, Wn, generator=g)
    a = torch.randn(Z, V, generator=g)
    ref = torch.einsum("zmu,uvw,zv->zmw", x.double().reshape(Z, d, U), W.double(), a.double()).reshape(Z * d, Wn)
    Bm = W.permute(2, 1, 0).reshape(Wn * V, U).contiguous()          # rows (w, v), v fastest
    C = torch.zeros(Z * d, Wn, device=DEV)
    ops.gemm_tf32x3(x.to(DEV), Bm.to(DEV), C, Z * d, Wn * V, U, reduce_aux=a.to(DEV), aux_d=d)
    assert _rel(C, ref) < 2e-6
