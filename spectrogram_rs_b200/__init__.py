"""Importable alias of the `spectrogram-rs_b200/` package directory (a hyphen cannot appear in a
Python module name).  All code lives in `spectrogram-rs_b200/`; this package only points there."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "spectrogram-rs_b200"))

from . import binding, sharding  # noqa: E402,F401
from .plan import MultiPlan, Plan, SbxError, builtin_gradient, builtin_gradient_names, builtin_lut, samples_from_seconds  # noqa: E402,F401
from .audio_transform import (AudioStreamTransform, AudioTransform, ColorScheme, FastFourierTransform,  # noqa: E402,F401
                              HeapRb, SpectrogramColumns, default_color_schemes)
