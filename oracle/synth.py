"""The seeded synthetic rectified-stereo generator lives in the package (front_end_b200/synth.py, numpy only) because
bench.py's GPU arm needs its inputs without touching oracle/; the oracle-side tests keep importing it from here."""
from front_end_b200.synth import *  # noqa: F401,F403
from front_end_b200.synth import MARGIN, _upsample4  # noqa: F401
