"""Index arithmetic of the per-species row order (`ops.species_row_groups`, used by the grouped tcgen05 GEMM of the
self-connection): a permutation of the rows sorted by species, every species padded to whole 128-row blocks, one weight
set per block.  Pure torch, runs on the CPU."""
import pytest
import torch

from e3b200 import ops


@pytest.mark.parametrize("N,S", [(1117, 7), (1, 3), (300, 20), (128, 2), (129, 1)])
def test_species_row_groups(N, S):
    g = torch.Generator().manual_seed(N + S)
    idx = torch.randint(0, S, (N,), generator=g)
    if S > 2:
        idx[idx == 1] = 0                               # an absent species takes no block
    grp = ops.species_row_groups(idx, S)
    assert grp.n_virtual == 128 * ((N + 127) // 128 + S) and grp.n_sets == S
    rm = grp.row_map
    assert rm.dtype == torch.int32 and rm.shape == (grp.n_virtual,) and grp.b_sel.shape == (grp.n_virtual // 128,)
    slots = torch.nonzero(rm >= 0).view(-1)
    assert sorted(rm[slots].tolist()) == list(range(N))                 # every row exactly once
    species = idx[rm[slots].long()]
    assert torch.equal(grp.b_sel[slots // 128].long(), species)         # its block carries its species' weight set
    assert bool((species[1:] >= species[:-1]).all())                    # ordered by species
    same = species[1:] == species[:-1]
    assert bool((rm[slots][1:][same] > rm[slots][:-1][same]).all())     # stable inside a species
    assert int(grp.b_sel.min()) >= 0 and int(grp.b_sel.max()) < S
    used = int(grp.n_blocks)                                            # the kernel stops after the blocks in use
    assert grp.n_blocks.dtype == torch.int32 and used == int(slots.max()) // 128 + 1
    assert bool((rm[used * 128:] < 0).all())
