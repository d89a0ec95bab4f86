"""CPU oracle package (test infrastructure only; see oracle/ferromic_oracle.h)."""
