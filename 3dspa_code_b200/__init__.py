"""3dspa_code_b200 - B200-native (sm_100a) implementation of the 3DSPA hot path.

Import with ``importlib.import_module("3dspa_code_b200")`` (the package name starts with a digit).
Importing loads lib3dspa_b200.so and fails loudly if it is missing: there is no CPU fallback.
"""
from . import _lib

_lib.lib()  # fail at import time if the CUDA library is not built

from . import data, lifting, ops, params  # noqa: E402
from .engine import DecoderContext, DeviceWeights, Engine  # noqa: E402
from .model import (  # noqa: E402
    TrackAutoEncoder,
    TrackAutoEncoder3D,
    TrackAutoEncoderDecoderContext,
    TrackAutoEncoderResults,
)
from .params import check_structure, load_checkpoint, save_checkpoint  # noqa: E402

__all__ = [
    "TrackAutoEncoder3D", "TrackAutoEncoder", "TrackAutoEncoderResults", "TrackAutoEncoderDecoderContext",
    "DecoderContext", "DeviceWeights", "Engine", "load_checkpoint", "save_checkpoint", "check_structure",
    "data", "lifting", "ops", "params",
]
