"""CPU oracle package -- test infrastructure only (see fb_oracle.py header)."""
