"""TEST INFRASTRUCTURE — not product code.

``oracle/`` holds the CPU restatement of the reference's planner hot path
(``planner_port.py``) and the shim that imports the real reference modules from
``/root/reference`` (``ref_import.py``, only usable in the authoring container).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this package, and only as the checker or the timed baseline.
Nothing under ``etpnav_b200/`` imports it.
"""
