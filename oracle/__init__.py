"""CPU oracle (test infrastructure only) -- see pme_oracle.py."""
