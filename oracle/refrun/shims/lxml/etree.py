"""Shim for the few `lxml.etree` entry points the reference's URDF loader touches."""
import xml.etree.ElementTree as _ET

Element = _ET.Element
ElementTree = _ET.ElementTree
tostring = _ET.tostring
fromstring = _ET.fromstring
SubElement = _ET.SubElement


class XMLParser:
    def __init__(self, remove_comments=False, remove_blank_text=False, **kw):
        pass


def parse(source, parser=None):
    return _ET.parse(source)
