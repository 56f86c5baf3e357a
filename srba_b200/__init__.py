"""srba_b200 — B200-native (sm_100a CUDA, FP64) rebuild of MRPT/srba's local Levenberg–Marquardt hot path.
See DESIGN.md.  The package holds only what the path needs: csrc/ (CUDA kernels + C ABI + host engine),
the ctypes binding, the Python mirror of RbaEngine<> and synthetic dataset generators."""
from . import _capi  # noqa: F401
from ._capi import (NOISE_CONST_MATRIX, NOISE_IDENTITY, OBS_CARTESIAN3D, OBS_MONO, OBS_RELPOSES2D, OBS_STEREO, POSE_SE2, POSE_SE3,  # noqa: F401
                    SENSORPOSE_NONE, SENSORPOSE_SE3, SOLVER_NO_SCHUR_SPARSE, SOLVER_SCHUR_DENSE, SOLVER_SCHUR_SPARSE, SrbaError, default_engine_params)
from .batch import Batch  # noqa: F401
from .engine import FlatProblem, RbaEngine, ResultBuffers, optimize  # noqa: F401
