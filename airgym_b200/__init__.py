"""airgym_b200 — B200-native drop-in for the batched-sim hot path of emNavi/AirGym."""
__version__ = "0.1.0"
