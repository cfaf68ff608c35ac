"""CPU oracle for the LSHM deep-K-harmonic hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lshm_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker or as the timed CPU arm.

Parity status: the reference ships no tests / golden vectors for this path
(SURVEY.md §4, §8c), so the oracle is pinned against the *live* reference code
imported from /root/reference in the authoring container
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``; ``tests/test_oracle_vs_reference.py``
re-checks directly whenever /root/reference is present).
"""
