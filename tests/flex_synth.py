"""Shim: the synthetic span store lives in the package (vrdd_b200.synth_flex) since bench.py uses it too."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vrdd_b200.synth_flex import *  # noqa: F401,F403,E402
from vrdd_b200.synth_flex import BINS, _integral  # noqa: F401,E402
