"""Shim: `lxml.etree` mapped onto the standard library ElementTree (lxml is absent from this image)."""
