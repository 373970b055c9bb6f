"""Mirror of the reference's ``gnn_model`` package (Neural Diving baseline on the shared bipartite GNN)."""
