"""CPU oracle (test infrastructure only): see oracle/oracle.py and oracle/bm25_oracle.c."""
