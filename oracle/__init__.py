"""TEST INFRASTRUCTURE ONLY.

CPU restatement ("oracle") of the DiT4SR denoising hot path.  Nothing in the product package
(`dit4sr_b200/`) may import this package; only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs use it, and only as the checker.
"""
