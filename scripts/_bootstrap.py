"""Put the in-tree `atmonr` package (atmospheric-neural-rendering_b200/atmonr) on sys.path."""

import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "atmospheric-neural-rendering_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
