"""CPU oracle of the FACT hot path (test infrastructure; see fact_oracle.py header)."""
