"""The seeded synthetic-input generator shared by the oracle, the parity tests and the bench lives in
``synthetic_data.py`` at the repo root (it is data generation, not reference arithmetic, and ``bench.py``'s GPU arm must
not import ``oracle/``); this module re-exports it under its historical name."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from synthetic_data import (_rng, one_hot, state_dict_spec, synth_image, synth_label, synth_pixel_weight,  # noqa: E402,F401
                            synth_state_dict)
