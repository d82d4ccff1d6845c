"""Locate and import the ctypes harness of the product library (freefem-sources_b200/ffcuda)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "freefem-sources_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

import ffcuda  # noqa: E402

HEADER = os.path.join(ROOT, "include", "ffcuda.h")
