"""uint8 host interface of the generator (``hv_pipeline_*`` in include/hv_b200.h).

The reference's inference driver feeds the generator uint8 planes and keeps uint8 results
(``eval_3d_sagittal_twostage.py:84-98`` in, ``:103-121`` out); ``SlicePipeline`` is that boundary for HOST buffers:

    pipe = SlicePipeline(generator, batch=16, depth=4)
    s = pipe.slot(0)                      # numpy views of the slot's PINNED host blocks
    s.ct[:n], s.cam[:n] = ct_u8, cam_u8   # [n, 256, 256] uint8 (CT already composed, CAM * 255)
    s.rows[:n] = (min_x, max_x + 1)       # mask rows [r0, r1) per slice (eval:73-75)
    s.ratio[:n] = index_ratio
    pipe.submit(0, n)                     # 1 H2D copy, ONE CUDA-graph launch of the two-stage forward, 1 D2H copy
    ...                                   # fill / submit the other slots meanwhile
    pipe.wait(0)                          # s.ct_out (uint8 CT), s.fine_mask / s.coarse_mask ({0,1}), s.heights [2, n]

Weights are read from the generator's native plan: call ``refresh()`` after changing them (training between evaluations).
There is no CPU fallback.
"""
import ctypes
import weakref
from collections import namedtuple

import numpy as np
import torch

from . import _lib
from ._lib import check

Slot = namedtuple("Slot", "ct cam rows ratio ct_out fine_mask coarse_mask heights")


class SlicePipeline:
    def __init__(self, generator, batch=16, depth=4, per_sample_mask=None, use_graph=True):
        self.g = generator
        self.batch, self.depth = int(batch), int(depth)
        dev = next(generator.parameters()).device
        if dev.type != "cuda":
            raise _lib.HvError("SlicePipeline needs the generator on a CUDA device (no CPU fallback)")
        if generator.training:
            raise _lib.HvError("SlicePipeline is the inference path: call generator.eval() first")
        self.dev = dev
        with torch.cuda.device(dev):
            plan = generator._ensure_plan(self.batch, dev)
            torch.cuda.current_stream(dev).synchronize()   # the plan was prepared on torch's stream, the pipeline runs on its own
            per_sample = generator.per_sample_mask if per_sample_mask is None else per_sample_mask
            handle = _lib.c_void_p()
            check(_lib.lib().hv_pipeline_create(ctypes.byref(handle), plan, self.batch, self.depth, int(bool(per_sample)),
                                                int(bool(use_graph))))
        self._h = handle
        self._plan = plan
        generator._pipelines.append(weakref.ref(self))
        self._slots = [self._views(i) for i in range(self.depth)]
        L = _lib.lib()
        self.h2d_bytes = int(L.hv_pipeline_bytes(self._h, 0))
        self.d2h_bytes = int(L.hv_pipeline_bytes(self._h, 1))

    def _views(self, i):
        ptrs = [_lib.c_void_p() for _ in range(8)]
        check(_lib.lib().hv_pipeline_slot(self._h, i, *[ctypes.byref(p) for p in ptrs]))
        b = self.batch

        def view(p, ctype, shape, dtype):
            n = int(np.prod(shape))
            return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctype)), shape=(n,)).view(dtype).reshape(shape)

        return Slot(view(ptrs[0], ctypes.c_uint8, (b, 256, 256), np.uint8), view(ptrs[1], ctypes.c_uint8, (b, 256, 256), np.uint8),
                    view(ptrs[2], ctypes.c_int32, (b, 2), np.int32), view(ptrs[3], ctypes.c_float, (b,), np.float32),
                    view(ptrs[4], ctypes.c_uint8, (b, 256, 256), np.uint8), view(ptrs[5], ctypes.c_uint8, (b, 256, 256), np.uint8),
                    view(ptrs[6], ctypes.c_uint8, (b, 256, 256), np.uint8), view(ptrs[7], ctypes.c_float, (2, b), np.float32))

    def slot(self, i):
        return self._slots[i]

    def stream(self, which):
        """The pipeline's CUDA streams as torch streams: 0 input copies, 1 compute, 2 output copies (event timing)."""
        return torch.cuda.ExternalStream(int(_lib.lib().hv_pipeline_stream(self._h, which)), device=self.dev)

    def refresh(self):
        """Re-prepare the plan if the generator's weights changed since the last forward / refresh.  Call with no slot in
        flight: the weights are re-packed on torch's current stream, which is synchronised before the pipeline's streams go on."""
        with torch.cuda.device(self.dev):
            plan = self.g._ensure_plan(self.batch, self.dev)
            torch.cuda.current_stream(self.dev).synchronize()
        if plan is not self._plan and plan.value != self._plan.value:
            raise _lib.HvError("the generator's native plan was rebuilt (precision / device / batch change): create a new SlicePipeline")

    def submit(self, i, n=None):
        check(_lib.lib().hv_pipeline_submit(self._h, i, self.batch if n is None else int(n)))

    def wait(self, i):
        check(_lib.lib().hv_pipeline_wait(self._h, i))

    def close(self):
        if self._h is not None:
            _lib.lib().hv_pipeline_destroy(self._h)
            self._h = None
            self._slots = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
