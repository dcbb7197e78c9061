"""CPU oracle for the STEm-Seg hot path (TEST INFRASTRUCTURE ONLY).

Nothing in ``stemseg_b200`` (the product) may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs use it,
and there only as the checker / the timed CPU baseline.

Pinning: the reference ships no tests or golden vectors, so every function here is pinned against
outputs of the reference itself, executed in the build container from /root/reference by the scripts in
``tests/golden/gen_*.py`` (committed together with the fixtures they wrote).
"""
