"""CPU parity oracle (test infrastructure only; see oracle/dg_oracle.c)."""
