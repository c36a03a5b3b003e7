"""CPU oracle and live-reference build for the lane-NMS path: test infrastructure only (see oracle/oracle.py)."""
