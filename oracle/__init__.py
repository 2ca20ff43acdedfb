"""CPU oracle (test infrastructure only -- see step_oracle.py / lidar_oracle.c headers)."""
