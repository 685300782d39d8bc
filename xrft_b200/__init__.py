"""xrft_b200 -- B200-native spectral engine behind the xrft API."""
__version__ = "0.1.0"
