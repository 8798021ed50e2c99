"""CPU oracle for the Lorenz Energy Cycle hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the
product package ``lorenzcycletoolkit_b200``; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it (as the checker / the timed CPU baseline).
"""
