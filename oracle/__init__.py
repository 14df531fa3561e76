"""CPU oracle for the ManipulaPy trajectory-and-dynamics hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``manipulapy_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do.  See ``oracle.c`` for the
restated algorithm and the reference file:line citations.
"""

from .oracle_lib import Oracle, build_oracle, load_oracle  # noqa: F401
