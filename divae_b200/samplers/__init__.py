"""Samplers of the RBM prior: same module and class names as models/samplers in the reference."""
