"""CPU oracle (test infrastructure only): restatement of the reference algorithm, pinned to reference goldens."""
