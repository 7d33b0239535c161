"""Parity oracle for the KADAL Kriging hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker / the timed CPU baseline.

* ``kriging_oracle``  – numpy/scipy CPU restatement of the reference algorithm
  (``likelihood_func.py``, ``kernelfunc.py``, ``prediction.py``, ``samplingplan.standardize``).
* ``ref_loader``      – imports the *live* reference from ``/root/reference`` (build
  container only) to validate the restatement and to generate ``tests/golden``.

Pinning: the reference ships no tests / golden vectors for this path, so
the restatement is pinned against outputs of the reference itself run in the build
container (``tests/golden/*.npz`` made by ``tests/golden/make_golden.py``) and against the
17-digit known-answer values of SURVEY.md Appendix C.
"""
