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


def _ssp(z):
    return torch.nn.functional.softplus(z) - 0.6931471805599453


def test_gemm_activation_epilogues_and_accumulate():
    """epilogue 2: c * ssp(alpha * acc); epilogue 3: alpha * acc * d/dz[c ssp](z) from the stored output;
    accumulate adds to C"""
    g = torch.Generator().manual_seed(5)
    M, N, K, c = 700, 64, 64, 1.8782
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    z = 0.125 * (A.double() @ B.double().T)
    ref2 = c * _ssp(z)
    C = torch.empty(M, N, device=DEV)
    ops.gemm_tf32x3(A.to(DEV), B.to(DEV), C, M, N, K, alpha=0.125, epilogue=2, act_cst=c)
    assert _rel(C, ref2) < 2e-6
    # backward through the activation, derivative taken from the stored forward output
    G = torch.randn(M, 96, generator=g)
    W = torch.randn(N, 96, generator=g)
    ref3 = 0.3 * (G.double() @ W.double().T) * (c * torch.sigmoid(z))
    D = torch.empty(M, N, device=DEV)
    ops.gemm_tf32x3(G.to(DEV), W.to(DEV), D, M, N, 96, alpha=0.3, epilogue=3, H=C, act_cst=c)
    assert _rel(D, ref3) < 5e-6
    base = torch.randn(M, N, generator=g)
    D2 = base.to(DEV)
    ops.gemm_tf32x3(G.to(DEV), W.to(DEV), D2, M, N, 96, alpha=0.3, accumulate=True)
    assert _rel(D2, base.double() + 0.3 * (G.double() @ W.double().T)) < 2e-6


@pytest.mark.parametrize("K,N", [(64, 64), (64, 320), (192, 64), (384, 320)])
def test_gemm_grouped_launch(K, N):
    """several independent problems (irreps blocks with their own weights) in one kernel"""
    g = torch.Generator().manual_seed(K + N)
    Z, D = 777, 9 * K
    X = torch.randn(Z, D, generator=g)
    Xd = X.to(DEV)
    dims = [1, 3, 5]
    offs = [0, K, 4 * K]
    Ws = [torch.randn(N, K, generator=g) for _ in dims]
    packed = ops.gemm_pack([(w.to(DEV), 0, K, 0, 1, 1, 0, N, K) for w in Ws])
    out = torch.zeros(Z, 9 * N, device=DEV)
    probs = []
    for d, off, pw in zip(dims, offs, packed):
        o_off = off // K * N
        probs.append(ops.gemm_problem(Xd, pw, out, Z * d, a_off=off, a_rows=(D, K, d), c_off=o_off, c_rows=(9 * N, N, d)))
    ops.gemm_run(probs)
    for d, off, w in zip(dims, offs, Ws):
        a = X[:, off:off + d * K].reshape(Z * d, K)
        ref = (a.double() @ w.double().T).reshape(Z, d * N)
        o_off = off // K * N
        assert _rel(out[:, o_off:o_off + d * N], ref) < 2e-6


def test_gemm_many_row_tiles_long_k():
    """more row tiles than CTAs and a long K (ring wrap-around, chained accumulation)"""
    g = torch.Generator().manual_seed(11)
    M, N, K = 40000, 64, 1920
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    C = torch.empty(M, N, device=DEV)
    ops.gemm_tf32x3(A.to(DEV), B.to(DEV), C, M, N, K)
    assert _rel(C, A.double() @ B.double().T) < 2e-6
    M, N, K = 30000, 1920, 64
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    C = torch.empty(M, N, device=DEV)
    ops.gemm_tf32x3(A.to(DEV), B.to(DEV), C, M, N, K)
    assert _rel(C, A.double() @ B.double().T) < 2e-6


@pytest.mark.parametrize("d,N,K", [(1, 64, 64), (3, 64, 64), (5, 64, 64), (1, 256, 64), (3, 64, 256), (5, 32, 32)])
def test_gemm_grouped_rows_weight_sets(d, N, K):
    """virtual row order by species + one packed weight set per species (the per-species self-connection): every node's
    d rows times ITS species' weight, written in place; nodes of absent species / padding slots never touch C"""
    g = torch.Generator().manual_seed(17 * d + N + K)
    Z, S, D_in, D_out = 1117, 7, 640 if K * d <= 640 else K * d, 1408
    species = torch.randint(0, S, (Z,), generator=g)
    species[species == 4] = 2                                    # one species absent
    X = torch.randn(Z, D_in, generator=g)
    W = torch.randn(S, N, K, generator=g)                        # set s: B[n, k]
    a_off, c_off = 64 if D_in >= 64 + K * d else 0, 128
    xa = X[:, a_off:a_off + d * K].reshape(Z, d, K).double()
    ref = 0.61 * torch.einsum("zdk,znk->zdn", xa, W.double()[species])          # [z, d, n]
    grp = ops.species_row_groups(species.to(DEV), S)
    assert grp.n_virtual == 128 * ((Z + 127) // 128 + S)
    rm = grp.row_map.cpu()
    assert sorted(rm[rm >= 0].tolist()) == list(range(Z))
    Wd = W.to(DEV).contiguous()
    (Bp,) = ops.gemm_pack([(Wd, 0, K, 0, 1, 1, 0, N, K, S, N * K)])
    C = torch.full((Z, D_out), 7.0, device=DEV)
    ops.gemm_run([ops.gemm_problem(X.to(DEV), Bp, C, grp.n_virtual * d, a_off=a_off, a_rows=(D_in, K, d), c_off=c_off,
                                   c_rows=(D_out, N, d), alpha=0.61, groups=grp)])
    out = C[:, c_off:c_off + d * N].reshape(Z, d, N)
    assert _rel(out, ref) < 2e-6
    assert bool((C[:, :c_off] == 7.0).all()) and bool((C[:, c_off + d * N:] == 7.0).all())
    # accumulate into what C holds
    ops.gemm_run([ops.gemm_problem(X.to(DEV), Bp, C, grp.n_virtual * d, a_off=a_off, a_rows=(D_in, K, d), c_off=c_off,
                                   c_rows=(D_out, N, d), alpha=0.61, groups=grp, accumulate=True)])
    assert _rel(C[:, c_off:c_off + d * N].reshape(Z, d, N), 2 * ref) < 2e-6


@pytest.mark.parametrize("M,N,K", [(120037, 64, 192), (60000, 128, 288), (113665, 64, 1920)])
def test_gemm_long_k_many_row_tiles(M, N, K):
    """K-long problems with several row tiles per CTA (the streaming A ring, both MMA issuers, chained accumulators):
    plain, accumulate and the activation-derivative epilogue, odd tile count, ragged last tile"""
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g, device=DEV)
    B = torch.randn(N, K, generator=g, device=DEV)
    ref = (0.5 * (A.double() @ B.double().T)).cpu()
    C = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm_tf32x3(A, B, C, M, N, K, alpha=0.5)
    assert _rel(C, ref) < 2e-6
    ops.gemm_tf32x3(A, B, C, M, N, K, alpha=0.5, accumulate=True)
    assert _rel(C, 2 * ref) < 2e-6
    cst = 1.3
    z = torch.randn(M, N, generator=g, device=DEV)
    H = cst * (torch.nn.functional.softplus(z) - 0.6931471805599453)
    ops.gemm_tf32x3(A, B, C, M, N, K, alpha=0.5, epilogue=3, H=H, act_cst=cst)
    ref3 = ref * (cst * torch.sigmoid(z.double())).cpu()
    assert _rel(C, ref3) < 4e-6


@pytest.mark.parametrize("M,N,K", [(70001, 1920, 64), (150011, 64, 64), (40003, 200, 8)])
def test_gemm_tensor_store_of_plain_outputs(M, N, K):
    """large plain row-major outputs of epilogues 0 / 2 leave through TMA tensor stores (one swizzled 32 x 32 block per
    epilogue warp and column step): ragged last row tile, N not a multiple of 32, untouched neighbours"""
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g, device=DEV)
    B = torch.randn(N, K, generator=g, device=DEV)
    acc = 0.25 * (A.double() @ B.double().T)
    pitch = N + 8
    buf = torch.full((M + 1, pitch), 3.0, device=DEV)
    C = buf[:M, :N]
    (Bp,) = ops.gemm_pack([(B, 0, K, 0, 1, 1, 0, N, K)])
    ops.gemm_run([ops.gemm_problem(A, Bp, buf, M, c_rows=(pitch, 0, 1), alpha=0.25)])
    assert _rel(C, acc.cpu()) < 2e-6
    assert bool((buf[:M, N:] == 3.0).all()) and bool((buf[M] == 3.0).all())
    cst = 1.7
    ops.gemm_run([ops.gemm_problem(A, Bp, buf, M, c_rows=(pitch, 0, 1), alpha=0.25, epilogue=2, act_cst=cst)])
    ref = cst * (torch.nn.functional.softplus(acc) - 0.6931471805599453)
    assert _rel(C, ref.cpu()) < 4e-6
    assert bool((buf[:M, N:] == 3.0).all()) and bool((buf[M] == 3.0).all())
