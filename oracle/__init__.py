"""TEST INFRASTRUCTURE ONLY.  CPU restatements of the reference's hot-path algorithms, used as the parity oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package;
the product (vistaocr_b200/) never does and has no CPU fallback.
"""
