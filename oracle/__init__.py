"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the SAR hot path.

Nothing in the product package (``nis-sar-amtigmti-video_b200/``) may import this
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
the timed CPU baseline -- never as the thing that produces a shipped result.

Parity status: **pinned against the reference itself run in the build
container** (the reference ships no golden vectors or tests of its own,
SURVEY.md section 8c).  ``oracle/make_golden.py`` AST-extracts the reference's
own functions from ``/root/reference`` (never copied into this repo), runs them
on seeded inputs and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` holds this restatement to those vectors.
"""
