"""B200-native counterpart of the reference's ``src.flux`` package (lightweight FLUX DiT + autoencoder)."""
