"""TEST INFRASTRUCTURE: parity oracle for the surfel reconstruction hot path.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this package; the product (surfelmeshing_b200/) never does.
"""
