"""Alias of the workload generator (mmser_b200/synth.py): tests and the golden generator import it from here."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mmser_b200.synth import *  # noqa: E402,F401,F403
from mmser_b200.synth import _fill, _gen  # noqa: E402,F401
