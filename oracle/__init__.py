"""CPU oracle for the closestMolecule FEM hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain numpy/scipy restatement of
what FEniCS 2019.1 (DOLFIN + FFC/uflacs + FIAT + PETSc LU) computes for the
reference scripts under /root/reference.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it -- and only as the checker or the timed CPU
baseline, never on the product path (``closestmolecule_b200`` never imports it
and fails loudly when its CUDA library is missing).

Third-party arithmetic restated here: legacy FEniCS 2019.1.0 ("tested with
v.2019.1", README.md:2 of the reference; not pinned by any manifest there).
FEniCS cannot run in this image, so the oracle is pinned on the reference's
recorded outputs instead:

* README.md:109-115   -- scalar P1 nonlinear Newton table (oracle.scalar, F-B):
  reproduced digit for digit by ``tests/test_oracle_golden.py``.
* convergence/test_for_paper_convergence.py:200-207 -- k=0 mixed table that
  exercises the P1 interior-facet/boundary-facet q-terms (oracle.paper_m0).

The P-Q primal P1 path itself (advection_diffusion_convergence.py:196-207) has
no recorded output anywhere in the reference: **parity unpinned** for it beyond
those two tables, finite-difference Jacobian checks and manufactured-solution
convergence rates.
"""
