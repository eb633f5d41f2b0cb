"""TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is the checker for the CUDA hot path: a CPU restatement
of the reference's algorithm (oracle/port.py, oracle/upstream.py), a harness that runs
the reference's own Python in place where ``/root/reference`` exists
(oracle/harness.py) and the scripts that generate tests/golden/.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs
may import it.  The product package ``robustsq_whisper_b200`` never does.
"""
