"""CPU oracle for the CMLPL hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``cmlpl_b200/`` may import this package.  It is imported by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and only as the checker / the CPU arm that is timed
beside the CUDA path -- never as the product.
"""
