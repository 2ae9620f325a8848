"""Importable alias for the package directory
``learning-temporal-consistency-for-video-scene-graph-generation_b200/`` (its name is not a valid
Python identifier).  ``import b200vsgg`` resolves sub-modules from that directory."""
import os as _os

_PKG_DIR = _os.path.normpath(_os.path.join(
    _os.path.dirname(_os.path.abspath(__file__)), "..",
    "learning-temporal-consistency-for-video-scene-graph-generation_b200"))
__path__.append(_PKG_DIR)
PKG_DIR = _PKG_DIR
