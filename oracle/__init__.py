"""ORACLE — test infrastructure only.

CPU fp32 restatement of the reference's composition hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (mvoc_b200/) never does.
"""
