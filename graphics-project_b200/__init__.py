"""graphics-project_b200: B200-native (sm_100a) line-detection engine.

Drop-in for the OpenCV edge/line front end of pklito/Graphics-Project:
``cv_shim`` mirrors the cv2 calls, ``detectors`` / ``vp`` mirror the reference's own wrappers,
``Engine`` is the batched CUDA-tensor API and ``dist`` shards batches by frame across GPUs.
Importing this package loads liblineengine.so and fails loudly if it is missing.
"""
from . import _native  # noqa: F401  (raises ImportError when the CUDA library is not built)
from .engine import Engine, EngineError, default_engine  # noqa: F401
from . import cv_shim, detectors, vp, dist, synth, pipeline  # noqa: F401

__version__ = _native.lib.le_version().decode()
