"""CPU oracle for the esloco hot path -- TEST INFRASTRUCTURE ONLY (see esloco_oracle.c)."""
