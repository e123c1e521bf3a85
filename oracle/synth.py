"""Deterministic synthetic weights / inputs / targets.

The generators live in the product package (``salt_b200.synthetic``: bench.py and the drop-in model use them
without touching ``oracle/``); the oracle re-exports them so that checker and engine see identical data.
"""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'open-solution-salt-identification_b200')
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from salt_b200.synthetic import (MEAN, STD, adapt_tiles, param_specs, resnet_block_counts, synth_inputs,  # noqa: E402,F401
                                 synth_salt_scenes, synth_state_dict, synth_targets, synth_tiles_u8)
