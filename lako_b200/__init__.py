"""lako_b200: B200-native (sm_100a) implementation of the LaKo FiDT5 reader hot path."""
__version__ = "0.1.0"
