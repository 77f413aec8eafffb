"""Synthetic calibration + event generator (test/bench infrastructure; see synth/npswf_synth.h)."""
from .synth import *  # noqa: F401,F403
