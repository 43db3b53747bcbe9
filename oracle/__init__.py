"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU checkers for the slice path).

Nothing under distributedllm_b200/ may import this package.  Allowed importers:
tests/, __graft_entry__.smoke(), bench.py (cpu_baseline leg and --impl reference).
"""
