"""Staged batch API (include/srba_b200.h): pack -> one H2D copy -> persistent LM kernel -> one D2H copy."""
import ctypes as C

import numpy as np

from . import _capi as K
from ._capi import lib
from .engine import ResultBuffers


class Batch:
    def __init__(self, flats, trace_cap=0, want_dense=False):
        self.L = lib()
        self.flats = list(flats)
        n = len(self.flats)
        arr = (C.POINTER(K.Problem) * n)(*[f.problem for f in self.flats])
        tc = -max(trace_cap, 1) if want_dense else trace_cap
        self.trace_cap, self.want_dense = trace_cap, want_dense
        self.h = self.L.srba_b200_batch_create(n, arr, tc)
        if not self.h:
            raise K.SrbaError(-1, self.L.srba_b200_last_error().decode())
        self.n = n

    def __del__(self):
        try:
            if self.h:
                self.L.srba_b200_batch_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _chk(self, rc):
        if rc != K.SRBA_OK:
            raise K.SrbaError(rc, self.L.srba_b200_last_error().decode())

    def upload(self, stream=0):
        self._chk(self.L.srba_b200_batch_upload(self.h, C.c_void_p(stream)))

    def update_values(self, i, edge_pose=None, lm_pos=None, obs_z=None):
        """srba_b200_batch_update_values: new numeric inputs for problem i (same structure); None keeps an array."""
        def p(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            self._keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_double))
        self._keep = []
        self._chk(self.L.srba_b200_batch_update_values(self.h, i, p(edge_pose), p(lm_pos), p(obs_z)))

    def upload_values(self, stream=0):
        self._chk(self.L.srba_b200_batch_upload_values(self.h, C.c_void_p(stream)))

    def run(self, stream=0):
        self._chk(self.L.srba_b200_batch_run(self.h, C.c_void_p(stream)))

    def download(self, stream=0):
        self._chk(self.L.srba_b200_batch_download(self.h, C.c_void_p(stream)))

    def unpack(self, i):
        res = ResultBuffers(self.flats[i].p, self.trace_cap, self.want_dense)
        self._chk(self.L.srba_b200_batch_unpack(self.h, i, C.byref(res.r)))
        return res

    def plan(self, i):
        """srba_b200_batch_problem_plan: [image bytes, workspace bytes, result bytes, shared memory, solver + 2 cluster + 4 wave,
        image arrays on chip, workspace arrays on chip, shared-memory budget] of problem i as this batch planned it."""
        out = (C.c_int64 * 8)()
        self._chk(self.L.srba_b200_batch_problem_plan(self.h, i, out))
        return list(out)

    def stats(self, i):
        r = K.Result()
        self.L.srba_b200_batch_unpack(self.h, i, C.byref(r))
        return r.stats.as_dict()

    STAGES = ("init", "schur_fill", "cholesky", "backsub", "features", "update", "chain", "residuals", "rho", "jacobians", "hessian",
              "gradient", "control", "output")

    SOLVER_SUBSTAGES = ("chol.gather", "chol.factor_loop", "chol.backsub_loop", "chol.w0_W", "chol.w0_bar1", "chol.w0_U", "chol.w0_bar2", "chol.iwarp_UI")

    def stage_cycles(self, i):
        out = np.zeros(24, np.uint64)
        self._chk(self.L.srba_b200_batch_stage_cycles(self.h, i, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        d = dict(zip(self.STAGES, out[:len(self.STAGES)].tolist()))
        d.update(zip(self.SOLVER_SUBSTAGES, out[16:16 + len(self.SOLVER_SUBSTAGES)].tolist()))
        return d

    @property
    def h2d_bytes(self):
        return self.L.srba_b200_batch_h2d_bytes(self.h)

    @property
    def d2h_bytes(self):
        return self.L.srba_b200_batch_d2h_bytes(self.h)

    @property
    def kernel_launches(self):
        return self.L.srba_b200_batch_kernel_launches(self.h)

    def algorithmic_bytes(self):
        out = np.zeros(5)
        self.L.srba_b200_batch_algorithmic_bytes(self.h, out.ctypes.data_as(C.POINTER(C.c_double)))
        return dict(B_chain=out[0], B_assem=out[1], B_schur=out[2], B_chol=out[3], F_chol=out[4])
