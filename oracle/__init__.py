"""CPU oracle -- test infrastructure only (see oracle/ref_port.py header)."""
