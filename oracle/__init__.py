"""TEST INFRASTRUCTURE ONLY: CPU checkers for the muon DCS hot path.

`oracle.port`      - ctypes binding of oracle/dcs_oracle.c (plain-C restatement).
`oracle.reference` - ctypes binding of oracle/_ref/libnoa_ref.so (the unmodified reference
                     headers compiled by oracle/Makefile), or None where it was never built.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  noa_b200/ (the product) must never do so.
"""
from .cpu import (Checker, build_port, build_reference, integral_scalar, load_port,  # noqa: F401
                  load_reference, material_assembly)
