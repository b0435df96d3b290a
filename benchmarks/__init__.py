"""Measurement scripts (run on a B200 box); results are committed under profiles/."""
