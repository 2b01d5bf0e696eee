"""streamoptima_b200 -- B200-native (sm_100a) implementation of StreamOptima's per-block encode hot path."""
__version__ = "0.1.0"
