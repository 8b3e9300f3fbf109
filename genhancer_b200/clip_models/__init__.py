"""B200-native counterpart of the reference's ``clip_models`` package."""
