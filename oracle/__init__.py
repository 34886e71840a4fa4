"""CPU oracle for the fedoo global-operator assembly path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fedoo_b200/`` may import this
package: it is the checker for the CUDA path (``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``), never the thing measured or shipped.
"""
