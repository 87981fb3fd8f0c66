"""Seeded synthetic workloads W1-W5.  Pure torch on CPU; callers move
the tensors to the GPU.  No reference data is needed (there is no network for datasets)."""
import math

import torch


def _molecule(n, gen, min_sep=1.0):
    """n points rejection-sampled uniformly in a ball of radius 0.85 n^(1/3) + 0.6 A with a
    minimum pairwise separation."""
    R = 0.85 * n ** (1.0 / 3.0) + 0.6
    pts = torch.zeros(n, 3, dtype=torch.float64)
    k = 0
    tries = 0
    while k < n:
        tries += 1
        p = (torch.rand(3, generator=gen, dtype=torch.float64) * 2 - 1) * R
        if p.norm() > R:
            continue
        if k == 0 or (pts[:k] - p).norm(dim=1).min() >= min_sep or tries > 2000 * n:
            pts[k] = p
            k += 1
    return pts


def qm9_like(n_graphs, seed=0, n_min=3, n_max=29, species_choices=(1, 6, 7, 8, 9), dtype=torch.float32):
    """W1/W2/W3: QM9-shaped molecules.  Returns dict(pos [N,3], species [N,1] int64,
    _n_nodes [G,1] int64)."""
    gen = torch.Generator().manual_seed(seed)
    ns = torch.randint(n_min, n_max + 1, (n_graphs,), generator=gen)
    pos = torch.cat([_molecule(int(n), gen) for n in ns]).to(dtype)
    choices = torch.tensor(species_choices, dtype=torch.long)
    species = choices[torch.randint(0, len(choices), (int(ns.sum()),), generator=gen)].view(-1, 1)
    return {"pos": pos, "species": species, "_n_nodes": ns.view(-1, 1).long()}


def diffusion_like(n_graphs, seed=0, n_min=3, n_max=29, num_types=18, dtype=torch.float32):
    """W4: pos = randn, complete graphs, bond_type ~ U{0..3} per edge, t ~ U(1e-5, 1) per graph."""
    gen = torch.Generator().manual_seed(seed)
    ns = torch.randint(n_min, n_max + 1, (n_graphs,), generator=gen)
    N = int(ns.sum())
    pos = torch.randn(N, 3, generator=gen, dtype=torch.float64).to(dtype)
    species = torch.randint(1, num_types, (N, 1), generator=gen)
    t = (torch.rand(n_graphs, 1, generator=gen, dtype=torch.float64) * (1 - 1e-5) + 1e-5).to(dtype)
    src, dst, off = [], [], 0
    for n in ns.tolist():
        ar = torch.arange(off, off + n)
        a, b = ar.repeat_interleave(n), ar.repeat(n)
        keep = a != b
        src.append(a[keep])
        dst.append(b[keep])
        off += n
    ei = torch.stack([torch.cat(src), torch.cat(dst)])
    bond = torch.randint(0, 4, (ei.shape[1], 1), generator=gen)
    n_edges = (ns * (ns - 1)).view(-1, 1).long()
    return {"pos": pos, "species": species, "t": t, "_n_nodes": ns.view(-1, 1).long(),
            "edge_index": ei, "bond_type": bond, "_n_edges": n_edges}


def protein_like(n_res, seed=0, n_chains=4, std=25.83, p_random=0.02, r_cut=8.0, dtype=torch.float32):
    """W5: one C-alpha graph: uniform in a ball at 1 residue / 135 A^3, scaled by 1/std;
    edges = (d < r_cut) U (same chain & |i-j| < 5) U Bernoulli(p_random), from a seeded mask."""
    gen = torch.Generator().manual_seed(seed)
    R = (n_res * 135.0 * 3.0 / (4.0 * math.pi)) ** (1.0 / 3.0)
    pts = []
    while len(pts) < n_res:
        p = (torch.rand(4 * n_res, 3, generator=gen, dtype=torch.float64) * 2 - 1) * R
        p = p[p.norm(dim=1) <= R]
        pts.extend(p[: n_res - len(pts)])
    ca = (torch.stack(pts) / std).to(dtype)
    idx = torch.arange(n_res)
    chain = (idx * n_chains // n_res).view(-1, 1)
    species = torch.randint(0, 21, (n_res, 1), generator=gen)
    a, b = idx.repeat_interleave(n_res), idx.repeat(n_res)
    d = torch.linalg.norm(ca[a] - ca[b], dim=-1)
    m = d < (r_cut / std)
    m |= (chain[a, 0] == chain[b, 0]) & ((a - b).abs() < 5)
    m |= torch.rand(n_res * n_res, generator=gen) < p_random
    m &= a != b
    ei = torch.stack([a[m], b[m]])
    t = torch.rand(1, 1, generator=gen, dtype=torch.float64).to(dtype) * (1 - 1e-5) + 1e-5
    return {"CA": ca, "species": species, "chain_id": chain, "id": idx.view(-1, 1), "t": t,
            "_n_nodes": torch.tensor([[n_res]]), "edge_index": ei,
            "_n_edges": torch.tensor([[ei.shape[1]]])}
