"""simmodel_b200: B200-native s2v / GNN-embedding hot path of rvdweerd/simmodel."""
