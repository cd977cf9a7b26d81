"""CPU oracle (test infrastructure only). See oracle/tgoracle.h."""
