"""google-nerf_b200: B200-native Instant-NGP hot path behind ngp_pl's NGP / render() API."""
__version__ = "0.1.0"
