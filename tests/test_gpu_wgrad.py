"""The split-K tcgen05 weight-gradient kernel (csrc/wgrad_tf32x3.cu, through the C ABI) against fp64 matmuls:
plain x^T g, affine row addressing of irreps blocks, the (feature x attribute) expansion of the self-connection
weights with strided output, grouped launches, ragged sizes; determinism (bitwise repeatable)."""
import pytest
import torch

from e3b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _err(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-300))


@pytest.mark.parametrize("R,K1,K2", [(1000, 64, 64), (149452, 64, 1920), (33, 8, 64), (5000, 128, 256), (70001, 64, 132),
                                     (31, 4, 4), (4096, 20, 36)])
def test_plain_xt_g(R, K1, K2):
    gen = torch.Generator().manual_seed(R + K1)
    x = torch.randn(R, K1, generator=gen).to(DEV)
    g = torch.randn(R, K2, generator=gen).to(DEV)
    out = ops.k_wgrad(x, g, alpha=0.37)
    ref = 0.37 * (x.double().t() @ g.double())
    assert out.shape == (K1, K2)
    assert _err(out, ref) < 5e-6, _err(out, ref)      # fp32 sums over up to 1.5e5 rows (chains of 256 rows, then split-K)
    assert torch.equal(out, ops.k_wgrad(x, g, alpha=0.37))          # deterministic: fixed-order split-K reduction


def test_irreps_blocks_grouped_with_accumulate():
    """rows = (node, component) of an irreps block inside a wider feature row, several blocks in one launch"""
    N, mul = 777, 64
    dims = [1, 3, 5]
    D = sum(d * mul for d in dims)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(N, D, generator=gen).to(DEV)
    g = torch.randn(N, D, generator=gen).to(DEV)
    W = torch.randn(3 * mul * mul, generator=gen).to(DEV)
    W0 = W.clone()
    probs, off, wo = [], 0, 0
    for d in dims:
        probs.append(ops.wgrad_problem(x, g, W, N * d, mul, mul, a_off=off, a_rows=(D, mul, d), b_off=off, b_rows=(D, mul, d),
                                       c_off=wo, alpha=0.5, accumulate=True))
        off += d * mul
        wo += mul * mul
    ops.wgrad_run(probs, x.device)
    off = wo = 0
    for d in dims:
        xb = x[:, off:off + d * mul].reshape(-1, mul).double()
        gb = g[:, off:off + d * mul].reshape(-1, mul).double()
        ref = W0[wo:wo + mul * mul].double().view(mul, mul) + 0.5 * xb.t() @ gb
        assert _err(W[wo:wo + mul * mul].view(mul, mul), ref) < 2e-6
        off += d * mul
        wo += mul * mul


@pytest.mark.parametrize("V,d", [(16, 1), (5, 3), (32, 5)])
def test_self_connection_weight_gradient(V, d):
    """dW[u, v, w] = alpha sum_{z, c} x[z, c, u] a[z, v] g[z, c, w], written straight into the [u, v, w] layout"""
    N, m1, mo = 901, 64, 32
    gen = torch.Generator().manual_seed(V)
    x = torch.randn(N, d * m1, generator=gen).to(DEV)
    g = torch.randn(N, d * mo, generator=gen).to(DEV)
    a = torch.randn(N, V, generator=gen).to(DEV)
    W = torch.full((m1 * V * mo,), 7.0, device=DEV)
    p = ops.wgrad_problem(x, g, W, N * d, m1, mo, a_rows=(d * m1, m1, d), b_rows=(d * mo, mo, d), aux=a, aux_d=d,
                          c_rows=(mo, V * mo, m1), alpha=1.3)          # row m = v * m1 + u  ->  u * V * mo + v * mo
    ops.wgrad_run([p], x.device)
    ref = 1.3 * torch.einsum("zcu,zv,zcw->uvw", x.double().view(N, d, m1), a.double(), g.double().view(N, d, mo))
    assert _err(W.view(m1, V, mo), ref) < 2e-6


def test_unsupported_widths_are_refused():
    x = torch.randn(10, 6, device=DEV)
    g = torch.randn(10, 8, device=DEV)
    with pytest.raises(RuntimeError):
        ops.k_wgrad(x, g)
