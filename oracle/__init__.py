"""ORACLE — test infrastructure only (see oracle/ddpm_oracle.py header)."""
