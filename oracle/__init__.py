"""CPU oracle for the tabular closed-loop hot path of CoBeL-RL.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline -- never as a code path of ``cobel_rl_b200``.

Contents
--------
philox.py      host (NumPy) definition of the per-agent uniform random stream
stream_rng.py  duck-typed ``numpy.random.Generator`` fed from a pre-drawn stream
ref_loader.py  imports the *real* reference from /root/reference (build
               container only) behind GUI/gym stubs; used to pin the oracle and
               to generate tests/golden/*.npz (see make_golden.py)
tabular.py     NumPy/Python restatement of the reference algorithms, every
               function citing the reference file:line it follows

Parity status: PINNED.  Every restated loop is checked bit-for-bit against the
reference itself run in this container (tests/test_oracle_vs_reference.py) and
against the committed golden vectors generated from the reference
(tests/golden/, generator: oracle/make_golden.py).  The reference's own
known-answer tests for this path (unit_tests/test_gridworld.py:28-41,
unit_tests/test_topology.py:93-109) are restated in tests/test_env_known_answers.py.
"""
