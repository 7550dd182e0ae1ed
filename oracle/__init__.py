"""CPU oracle (test infrastructure only) -- see oracle/triangulation_oracle.py and DESIGN.md."""
