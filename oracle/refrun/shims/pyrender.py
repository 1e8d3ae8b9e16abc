"""Shim: rendering is outside the inference path."""
