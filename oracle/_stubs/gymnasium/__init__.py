"""Minimal stand-in for `gymnasium` so the read-only reference imports headlessly.

Test infrastructure only (see oracle/__init__.py). Never imported by the product.
"""
from . import spaces
from .spaces import Space


class Env:
    pass
