"""CPU oracle for the OpenViVQA attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  The product path (``openvivqa_b200``) never imports it and
fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md §4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF,
run in the authoring container by ``oracle/make_golden.py`` (imports
``/root/reference`` through ``oracle/ref_loader.py``) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks the restatement
against those vectors everywhere, and against the live reference wherever
``/root/reference`` is mounted.
"""
