"""CPU oracle for the lfvio hot paths — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
anything from here, and only as the checker or the timed CPU baseline.  Nothing under
lf-vio_b200/ imports it; the product path has no CPU fallback.
"""
