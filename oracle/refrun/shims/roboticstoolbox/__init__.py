"""Shim: only used by the reference's renderer helper object."""
