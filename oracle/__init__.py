"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference hot path (martinambrus/PitchExtractor), used as
the checker by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  Nothing under ``pitchextractor_b200/``
may import this package: the product path is CUDA-only and fails loudly without
its extension.

Pinning status: the reference ships no tests / golden vectors, so
the oracle is pinned against the LIVE reference imported from ``/root/reference``
in the build container (``oracle/refshim.py`` + ``tests/test_oracle_vs_reference.py``)
and against fixtures generated from it (``tests/golden/make_golden.py``).
"""
