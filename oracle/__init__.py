"""CPU oracle -- TEST INFRASTRUCTURE ONLY.  See oracle/orb_oracle.h for who may import this."""
