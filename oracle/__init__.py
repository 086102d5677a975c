"""CPU oracle for the bayesian-ode hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a float64 CPU restatement of the reference's
algorithms (jaivardhankapoor/bayesian-ode; citations are file:line relative to
the reference checkout).  It exists to *check* the CUDA product path:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import it;
* nothing in ``bayesian-ode_b200/`` imports it, and the product has no CPU
  fallback (it raises when the CUDA library is missing).

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified
Python reference in the build container and stores its outputs as fixtures
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every oracle
function against those fixtures (and against the closed-form known-answer
problems of the reference's own ``neuralode_tests/problems.py``).
"""
