"""Shim: the silhouette renderer is outside the inference path."""
