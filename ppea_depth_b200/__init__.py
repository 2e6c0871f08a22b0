"""B200-native view-synthesis loss path for PPEA-Depth."""
__version__ = "0.1.0"
