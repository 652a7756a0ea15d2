"""CPU checker for optable_b200 (test infrastructure only; never imported by the product)."""
