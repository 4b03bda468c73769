"""more3d_b200: B200-native per-point neighbourhood-feature path of MorE3D."""
