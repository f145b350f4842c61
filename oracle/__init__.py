"""CPU oracle for the sts per-bitstream test kernels.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product package (sts_b200) never does.
"""
