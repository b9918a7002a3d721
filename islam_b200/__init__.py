"""islam_b200 — B200-native pose-velocity-graph optimisation (iSLAM back-end hot path)."""
__version__ = "0.1.0"
