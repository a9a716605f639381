"""qdsp_b200 — B200-native (sm_100a) implementation of qdsp's data-parallel signal chain behind the
reference's `dsp::` block interface. Product code only: the CPU oracle lives in ../oracle and is
never imported from here."""
from . import lib  # noqa: F401
from .blocks import *  # noqa: F401,F403

__version__ = "0.1.0"
