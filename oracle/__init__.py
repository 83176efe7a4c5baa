"""oracle/ - TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's sampling hot path plus the stubs that let the reference's
own Python import in the build container.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import anything from here, and only as
the checker or the timed CPU baseline - never as part of the product path
(`thermodynamic_interpolation_b200/`), which fails loudly when its CUDA library is missing.
"""
