"""CPU oracle package (test infrastructure only; see ``oracle/yee.py`` header)."""
