"""CPU oracle for the yacrd detect path (test infrastructure only — never imported by yacrd_b200/)."""
