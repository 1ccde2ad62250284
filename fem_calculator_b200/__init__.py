"""femb200 — B200-native numerical hot path for FEM-calculator."""
