"""Import alias for the product package.

The product directory is ``betrayed-by-captions_b200/`` (the name the build contract
asks for); a hyphen is not importable, so this thin package extends its ``__path__`` to
that directory: ``import cgg_b200.head`` loads ``betrayed-by-captions_b200/head.py``.
"""
import os as _os

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
PRODUCT_DIR = _os.path.join(_ROOT, 'betrayed-by-captions_b200')
__path__.append(PRODUCT_DIR)
