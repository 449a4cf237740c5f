"""microdit-b200: a B200 (sm_100a) native implementation of the MicroDiT training hot path."""
__version__ = "0.1.0"
