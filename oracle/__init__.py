"""CPU oracle for the TRIAD max-mean similarity + InfoNCE hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``triad_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` do.
"""
from .oracle import *  # noqa: F401,F403
