"""Host-side invariants of the tensor-product kernel generator (csrc/gen_tp.py): the committed header is what the generator
emits today, the warp groups of every structure cover its input blocks exactly once within their register budget, every
structure lands in exactly one translation-unit part, and the 9-value halving butterfly of the backward kernels
(common.cuh::gsh_reduce_store9) leaves value i on the lane the store reads it from (emulated lane by lane)."""
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(HERE, "..", "equivariant-nn-zoo_b200")
sys.path.insert(0, PKG)
sys.path.insert(0, os.path.join(PKG, "csrc"))

import gen_tp  # noqa: E402
from e3b200.plan import generated_structures  # noqa: E402


def test_committed_header_is_current(tmp_path):
    out = tmp_path / "tp_generated.cuh"
    env = dict(os.environ, E3B_GEN_ONLY=",".join(str(i) for i in range(len(generated_structures()))), E3B_GEN_OUT=str(out))
    subprocess.check_call([sys.executable, os.path.join(PKG, "csrc", "gen_tp.py")], env=env)
    with open(os.path.join(PKG, "csrc", "tp_generated.cuh")) as f:
        committed = f.read()
    assert out.read_text() == committed, "run `python equivariant-nn-zoo_b200/csrc/gen_tp.py` and commit tp_generated.cuh"


def test_groups_partition_the_input_blocks():
    for st in generated_structures():
        n_in = len(st.irreps_in)
        acc = [sum(p.ir_out.dim for p in st.paths if p.i_in == b) for b in range(n_in)]
        for cap in (None, max(gen_tp.PAIRED_MAX_ACC, max(acc))):
            groups = gen_tp.make_groups(st, cap)
            flat = sorted(b for g in groups for b in g)
            assert flat == list(range(n_in))
            limit = cap or gen_tp.MAX_ACC_PER_GROUP
            assert all(sum(acc[b] for b in g) <= limit for g in groups)


def test_every_structure_is_in_one_part_and_the_table_lists_all():
    with open(os.path.join(PKG, "csrc", "tp_generated.cuh")) as f:
        text = f.read()
    n = len(generated_structures())
    opens = re.findall(r"#if !defined\(__CUDACC__\) \|\| !defined\(E3B_TP_PART\) \|\| E3B_TP_PART == (\d+)\n// ---- structure S(\d+):", text)
    assert sorted(int(s) for _, s in opens) == list(range(n))
    assert {int(p) for p, _ in opens} <= set(range(gen_tp.N_PARTS))
    assert f"static const int kNumGenEntries = {n};" in text
    for sid in range(n):
        assert f"void launch_tpf_S{sid}(const TpArgs<float>& a" in text and f"void launch_tpb_S{sid}(const TpArgs<float>& a" in text
    srcs = os.listdir(os.path.join(PKG, "csrc"))
    assert all(f"tp_fast_p{k}.cu" in srcs for k in range(1, gen_tp.N_PARTS))


def test_halving_butterfly_emulation():
    """lane-by-lane emulation of gsh_reduce_store9: after the exchanges over lane bits 4, 3, 2 and the plain steps over bits
    1, 0, lane 4 i holds the total of value i (i < 8)"""
    rng = np.random.default_rng(0)
    r = rng.standard_normal((32, 9))
    lanes = np.arange(32)
    b4, b3, b2 = (lanes & 16) != 0, (lanes & 8) != 0, (lanes & 4) != 0

    def xor(v, m):
        return v[lanes ^ m]

    t = np.stack([np.where(b4, r[:, k + 4], r[:, k]) + xor(np.where(b4, r[:, k], r[:, k + 4]), 16) for k in range(4)], 1)
    s2 = np.stack([np.where(b3, t[:, k + 2], t[:, k]) + xor(np.where(b3, t[:, k], t[:, k + 2]), 8) for k in range(2)], 1)
    q = np.where(b2, s2[:, 1], s2[:, 0]) + xor(np.where(b2, s2[:, 0], s2[:, 1]), 4)
    q = q + xor(q, 2)
    q = q + xor(q, 1)
    row = np.zeros(8)
    for lane in range(0, 32, 4):
        row[(lane >> 4) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = q[lane]
    np.testing.assert_allclose(row, r[:, :8].sum(0), rtol=1e-12, atol=1e-12)
