"""CPU oracle for the Equivariant-NN-Zoo hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``equivariant-nn-zoo_b200/``) never imports it and never falls back to it.

What it is
----------
A pure-``torch`` (CPU, fp32 or fp64) restatement of

* the arithmetic of the reference's third-party dependency ``e3nn==0.4.4``
  (``/root/reference/requirements.txt:27``) and ``torch-runstats==0.2.0``
  (``requirements.txt:145``), neither of which is vendored in the reference nor
  installable here (no network, not in the wheelhouse) -- ``oracle/e3nn_ops.py``,
  ``oracle/irreps.py``, ``oracle/wigner.py``;
* the reference's own composition of those operators along the hot path
  (``e3_layers/nn/*.py``, ``e3_layers/data/compute_edge.py``) in the reference
  dataflow (per-edge materialisation, per-edge ``o3.Linear`` *before* the
  scatter, ``scatter_add``) -- ``oracle/ref_layers.py``.

PARITY UNPINNED at the e3nn boundary: the reference ships no tests, golden
vectors or fixtures, and real e3nn cannot be imported in the build container.
What pins the oracle instead:

* closed-form known-answer tests in ``tests/test_oracle_kat.py``;
* SO(3) rotation + inversion equivariance in fp64;
* the reference's OWN python modules (``/root/reference/e3_layers``) executed in
  the build container on top of ``oracle/shims.py`` (which routes ``import e3nn``
  to ``oracle/e3nn_ops.py``): ``tests/golden/make_golden.py`` generates the
  committed fixtures under ``tests/golden/`` that way, so the *composition*
  (layer order, irreps wiring, key mapping, normalisation constants applied by
  the reference's code) is pinned by the reference itself;
* ``tools/check_against_e3nn.py`` diffs every convention against genuine e3nn
  on any machine that has it.

Two convention risks remain: signs of real
``wigner_3j`` for odd ``l1+l2+l3`` and the Monte-Carlo ``normalize2mom``
constants.  Both are data tables in one place (``oracle/wigner.py``,
``oracle/e3nn_ops.py:NORMALIZE2MOM``).
"""
