"""CPU oracle for the bundle-adjustment hot path.  TEST INFRASTRUCTURE ONLY (see ba_oracle.c)."""
