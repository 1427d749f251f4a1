"""TEST INFRASTRUCTURE ONLY — CPU oracle of the TAV fusion hot path of g8a9/multi-modal-emotion.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or the timed CPU baseline — the
product path (``multi-modal-emotion_b200/``) never imports it and has no CPU fallback.

Parity pinning: the reference ships no tests for this path; the pins are (i) the two known-answer
values harvested from its notebooks (CE 1.9298 / 1.9443, conv-length f(3280)=10) and (ii) golden vectors produced by
importing and running the UNMODIFIED reference modules from /root/reference in the authoring container
(``oracle/make_golden.py`` -> ``tests/golden/*.pt``).  The restatement in ``oracle/tav_oracle.py`` is checked against
both in ``tests/test_oracle_cpu.py``.
"""
