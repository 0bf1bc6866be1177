"""TEST INFRASTRUCTURE ONLY: CPU oracle of the jaxngp hot path (see oracle/ngp_oracle.c header).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs.
"""
