"""Shim of torch_geometric (>= 1.6 API) -- only the names reference models/pytorch_geometric/pna.py:2,7-9 imports."""
