"""shapeclipper_b200 — B200 (sm_100a) implementation of ShapeClipper's hot path behind the reference's
own Python interfaces. Hand-written CUDA in csrc/ reached through the C ABI of include/sc_b200.h."""
__version__ = "0.1.0"
