"""vsdeoldify_b200 — B200-native (sm_100a) implementation of HAVC's per-frame colorization hot path."""
__version__ = "0.1.0"
