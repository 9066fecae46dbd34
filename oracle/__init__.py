"""ORACLE — test infrastructure.  CPU restatements of the reference's fused-op path used ONLY by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
