"""TEST INFRASTRUCTURE ONLY -- CPU oracles for the RAFT correlation hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package
(``rdvc_corr_b200``) never does and raises if its CUDA library is missing.

Three independent statements of the same arithmetic:

* :mod:`oracle.corr_numpy`   -- numpy restatement (einsum + explicit gather).
* :mod:`oracle.corr_c`       -- ctypes binding of ``corr_oracle.c`` (scalar C, OpenMP).
* :mod:`oracle.tv_corr`      -- the reference's actual dependency: torchvision's
  ``CorrBlock`` run on CPU (``kind: "reference"`` in bench.py's cpu_baseline).

Parity pinning: the reference repository holds no golden vectors for this
path ("parity unpinned" by the reference's own
tests).  The oracles are pinned instead against live torchvision 0.26.0
``CorrBlock`` runs and the fixtures in ``tests/golden/`` generated from it by
``tests/golden/make_golden.py``.
"""
