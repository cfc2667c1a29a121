"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the CTC beam-search hot path (see oracle/README.md).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (ctcdecode_b200/) never does.
"""
