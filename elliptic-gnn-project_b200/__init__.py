"""B200-native GNN message-passing hot path (GCN / SAGE / SAGE-ResBN / GAT)."""
