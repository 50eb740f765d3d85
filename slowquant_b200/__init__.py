"""B200-native state-vector engine for unitary product-state wave functions (SlowQuant call surface)."""
__version__ = "0.1.0"
