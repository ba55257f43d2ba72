"""B200-native OAR-TRANSEG -> DOSE-PYFER hot path."""
__version__ = "0.1.0"
