"""continual-skeletons_b200: B200-native continual ST-GCN per-step forward (import as
``continual_skeletons_b200``)."""
