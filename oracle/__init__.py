"""CPU oracle (TEST INFRASTRUCTURE ONLY; parity unpinned — see kanode_oracle.cpp header).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs.
"""
from .pyoracle import Oracle, build_oracle, ORACLE_LIB  # noqa: F401
