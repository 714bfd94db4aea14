"""Parity-checking infrastructure (CPU oracle + reference harnesses).
TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by heongpu_b200."""
