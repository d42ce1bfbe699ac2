"""CPU oracle for the NATriuM semi-Lagrangian stream + collide hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (numpy + plain C) of
the reference algorithm for the per-timestep hot path.  It exists to *check* the
CUDA product; nothing under ``natrium_b200/`` imports it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import, link or execute anything in here.

Parity pinning status
---------------------
* collide (BGK f-only, BGK f+g/quartic, legacy BGKStandard, entropic family): pinned by
  the reference's own known-answer / self-consistency tests restated in
  ``tests/test_oracle_*.py`` (BGKStandard_test.cpp, Equilibrium_test.cpp,
  KBCStandard_test.cpp -- see SURVEY.md section 4).
* stream (SpMV): the arithmetic lives in Trilinos 13.0.1 ``Epetra_CrsMatrix::Multiply``
  reached through deal.II 9.3.3 ``TrilinosWrappers::BlockSparseMatrix::vmult``; neither is
  vendored in /root/reference and the reference holds no element-wise golden vector for
  it, so **bitwise parity of the SpMV is unpinned**; it is pinned only through the
  reference's property tests (M*1 = 1, SemiLagrangian_test.cpp:519-596; uniform flow stays
  uniform, CFDSolver_test.cpp:44-112; config-1 E_kin decay, IntegrationTestCases.cpp:885-961).
* the reference itself cannot be built in this image (needs deal.II, Trilinos, p4est, Boost,
  MPI -- all absent), so there is no ``oracle/_ref``.
"""
