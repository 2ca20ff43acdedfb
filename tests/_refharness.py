"""Back-compat alias: the reference harness lives in oracle/ref_harness.py (test infrastructure)."""
from oracle.ref_harness import *          # noqa: F401,F403
from oracle.ref_harness import FixedNoise, available, load, make_models, make_refiner, make_trainer  # noqa: F401
