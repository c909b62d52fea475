"""CPU parity oracle -- test infrastructure only (see vds_oracle.c header)."""
