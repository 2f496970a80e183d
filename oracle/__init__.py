"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this package, and there only as the
checker or the timed CPU baseline — never as (part of) the product path.
See ``x360_oracle.c`` for what is restated and how parity is pinned.
"""
