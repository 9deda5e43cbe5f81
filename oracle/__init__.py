"""CPU oracle for the SLAF ML hot path -- test infrastructure only (see slaf_oracle.c header)."""
