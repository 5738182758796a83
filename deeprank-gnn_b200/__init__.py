"""B200-native DeepRank-GNN hot path."""
__version__ = "0.1.0"
