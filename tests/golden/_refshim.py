"""Thin alias of baseline/refshim.py (kept so that the golden generators' `import _refshim` keeps working)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from baseline.refshim import REFERENCE_ROOT, available, cpu_only, find_reference_root, install, load_config  # noqa: E402,F401
