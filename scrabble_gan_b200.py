"""Importable alias of the package directory `scrabble-gan_b200/` (a hyphen is not a valid identifier)."""
import importlib
import sys

_pkg = importlib.import_module("scrabble-gan_b200")
sys.modules[__name__] = _pkg
