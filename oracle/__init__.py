"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  megreader_b200/ never does.
"""
