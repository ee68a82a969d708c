"""CPU oracle for the Segtran hot path — TEST INFRASTRUCTURE ONLY.

Importable by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (segtran_b200/) never imports it.
"""
