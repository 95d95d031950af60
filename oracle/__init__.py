"""CPU oracle for the CMAD constitutive-update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import, link or execute it,
and there only as the checker / CPU baseline, never as the thing shipped.

Parity status: the reference (sandialabs/cmad) is pure Python on JAX, and JAX
is not installable in this environment, so the reference itself cannot be run
here.  The oracle is therefore pinned against the reference's own known-answer
constructions, all of which are formulas plus
tolerances restated in ``tests/``; parity with *actual JAX output* at 1e-10 is
unpinned ("parity unpinned" for that stronger claim).
"""
