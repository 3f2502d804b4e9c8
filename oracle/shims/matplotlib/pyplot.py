"""Import-only stub: reference graph_generation.py:5 imports pyplot but never calls it on the generation path."""
