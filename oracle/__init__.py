"""CPU oracle for the hot path -- TEST INFRASTRUCTURE ONLY (see lw_oracle.c header)."""
