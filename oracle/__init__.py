"""CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product package never does.
"""
from .oracle import Oracle, build_oracle, oracle_lib_path  # noqa: F401
