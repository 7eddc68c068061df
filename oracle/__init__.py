"""CPU oracle for the localrf render path -- TEST INFRASTRUCTURE ONLY (see lrf_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (localrf_b200/) never does.
"""
