"""CPU oracle for the MusicGAN hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``musicgan_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and there only as the checker / the CPU baseline.

The reference (Ipsedo/MusicGAN) is pure Python on torch; the restatements here
are plain torch-CPU fp32 code that follows the reference op by op (each function
cites the reference file:line).  They are pinned against the reference itself,
imported from /root/reference in the authoring container, by
``oracle/gen_golden.py`` which wrote the fixtures under ``tests/golden/``.
"""
