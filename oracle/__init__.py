"""CPU oracle for the SLODE latent-ODE hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``structured_latent_odes_b200`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  See ``oracle/README.md``.
"""
