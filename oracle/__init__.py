"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/haf_oracle.cpp header).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
