"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the GAOT-3D hot path (graph build, GNO, geometric embedding,
latent attention).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package, and only as the checker / the CPU arm -- never from the product
package ``gaot_3d_b200``.
"""
