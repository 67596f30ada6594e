"""cosim_b200: B200-native batched stepping engine for cosim's sim-to-sim evaluation hot path."""
__version__ = "0.1.0"
