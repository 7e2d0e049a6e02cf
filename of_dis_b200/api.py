"""Python host side of the C-ABI (include/ofdis_b200.h), via ctypes.

`Context` is the batch engine; `OFClass`, `PatGridClass`, `VarRefClass` mirror the
reference's three classes (oflow.h:84-111, patchgrid.h:19-44,
refine_variational.h:37-39) with the same argument meaning, so the parity tests
read like calls into the reference.  There is NO CPU fallback: importing this
module loads of_dis_b200/lib/libofdis_b200.so and raises if it is missing.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .params import CParams, DisParams

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OFDIS_LIB") or os.path.join(_HERE, "lib", "libofdis_b200.so")  # OFDIS_LIB: experiments only
_FP = ctypes.POINTER(ctypes.c_float)
_IP = ctypes.POINTER(ctypes.c_int)

MEM_HOST, MEM_DEVICE = 0, 1

EXPORTS = [
    "ofdis_create", "ofdis_destroy", "ofdis_last_error", "ofdis_version", "ofdis_level_info",
    "ofdis_upload_level", "ofdis_packed_frame_floats", "ofdis_packed_offset", "ofdis_upload_packed",
    "ofdis_patgrid_optimize", "ofdis_patgrid_aggregate", "ofdis_varref_refine", "ofdis_run", "ofdis_sync",
    "ofdis_get_flow", "ofdis_set_flow", "ofdis_get_flow_batch", "ofdis_get_patches", "ofdis_debug_get",
    "ofdis_debug_varref_iters", "ofdis_launch_count", "ofdis_set_graph_mode", "ofdis_profile_run",
    "ofdis_set_camlr", "ofdis_set_dp_thresh_sq", "ofdis_packed_images_frame_floats", "ofdis_upload_packed_images",
    "ofdis_upload_frames_u8", "ofdis_finest_level_frame_floats", "ofdis_upload_finest_level", "ofdis_get_flow_fullres",
    "ofdis_get_level", "ofdis_upload_level_fb", "ofdis_set_option", "ofdis_profile_levels", "ofdis_set_direction",
]


class OfdisError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads the CUDA library; raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OfdisError("%s not found: run `python -m of_dis_b200.build` (no CPU fallback exists)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.ofdis_last_error.restype = ctypes.c_char_p
        L.ofdis_version.restype = ctypes.c_char_p
        L.ofdis_packed_frame_floats.restype = ctypes.c_size_t
        L.ofdis_packed_images_frame_floats.restype = ctypes.c_size_t
        L.ofdis_packed_offset.restype = ctypes.c_size_t
        L.ofdis_debug_get.restype = ctypes.c_long
        L.ofdis_launch_count.restype = ctypes.c_long
        L.ofdis_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_void_p,
                                   ctypes.POINTER(CParams), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int]
        for name in ("ofdis_destroy", "ofdis_sync"):
            getattr(L, name).argtypes = [ctypes.c_void_p]
        L.ofdis_last_error.argtypes = [ctypes.c_void_p]
        L.ofdis_launch_count.argtypes = [ctypes.c_void_p]
        L.ofdis_packed_frame_floats.argtypes = [ctypes.c_void_p]
        L.ofdis_packed_images_frame_floats.argtypes = [ctypes.c_void_p]
        L.ofdis_upload_packed_images.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ofdis_packed_offset.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.ofdis_get_level.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ofdis_finest_level_frame_floats.restype = ctypes.c_size_t
        L.ofdis_finest_level_frame_floats.argtypes = [ctypes.c_void_p]
        L.ofdis_upload_finest_level.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ofdis_upload_frames_u8.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int]
        L.ofdis_get_flow_fullres.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int]
        L.ofdis_upload_packed.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ofdis_upload_level.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4 + [ctypes.c_int]
        L.ofdis_upload_level_fb.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.c_int]
        L.ofdis_level_info.argtypes = [ctypes.c_void_p, ctypes.c_int] + [_IP] * 5
        L.ofdis_patgrid_optimize.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4
        L.ofdis_patgrid_aggregate.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3
        L.ofdis_varref_refine.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3
        L.ofdis_debug_varref_iters.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4
        L.ofdis_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.ofdis_get_flow.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ofdis_set_flow.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ofdis_get_flow_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ofdis_get_patches.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4
        L.ofdis_debug_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        L.ofdis_set_graph_mode.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.ofdis_set_option.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        L.ofdis_set_direction.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.ofdis_set_camlr.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.ofdis_set_dp_thresh_sq.argtypes = [ctypes.c_void_p, ctypes.c_float]
        L.ofdis_profile_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.ofdis_profile_levels.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(int(a))  # raw address (e.g. torch tensor .data_ptr())


class Context:
    """One device, one stream, frames [0, max_frames) per launch."""

    def __init__(self, prm: DisParams, width: int, height: int, imgpadding: int | None = None, max_frames: int = 1,
                 device: int = 0, stream: int | None = None):
        self.prm = prm
        self.width, self.height = width, height
        self.pad = prm.p_samp_s if imgpadding is None else imgpadding
        self.max_frames = max_frames
        self._h = ctypes.c_void_p()
        cp = prm.to_c()
        rc = lib().ofdis_create(ctypes.byref(self._h), device, ctypes.c_void_p(stream or 0), ctypes.byref(cp), prm.nop,
                                width, height, self.pad, max_frames)
        if rc != 0:
            self._h = ctypes.c_void_p()
            raise OfdisError("ofdis_create failed with status %d" % rc)

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if self._h:
            lib().ofdis_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise OfdisError("status %d: %s" % (rc, lib().ofdis_last_error(self._h).decode()))

    # -- geometry ----------------------------------------------------------
    def level_info(self, level: int):
        v = [ctypes.c_int() for _ in range(5)]
        self._ck(lib().ofdis_level_info(self._h, level, *[ctypes.byref(x) for x in v]))
        return dict(zip(("w", "h", "nopw", "noph", "steps"), [x.value for x in v]))

    @property
    def packed_frame_floats(self) -> int:
        return lib().ofdis_packed_frame_floats(self._h)

    def packed_offset(self, level: int, which: int) -> int:
        return lib().ofdis_packed_offset(self._h, level, which)

    def pack_frame(self, pyr, out: np.ndarray | None = None) -> np.ndarray:
        """Lays one PairPyramids out in the context's packed transfer format."""
        buf = np.zeros(self.packed_frame_floats, np.float32) if out is None else out
        for lv in range(self.prm.sc_l, self.prm.sc_f + 1):
            for k, arr in enumerate((pyr.i0[lv], pyr.i0x[lv], pyr.i0y[lv], pyr.i1[lv])):
                o = self.packed_offset(lv, k)
                buf[o:o + arr.size] = arr.reshape(-1)
        return buf

    # -- transfers ----------------------------------------------------------
    def upload_level(self, frame, level, i0, i0x, i0y, i1, memkind=MEM_HOST):
        self._ck(lib().ofdis_upload_level(self._h, frame, level, _ptr(i0), _ptr(i0x), _ptr(i0y), _ptr(i1), memkind))

    def upload_level_fb(self, frame, level, i0, i0x, i0y, i1, i1x, i1y, memkind=MEM_HOST):
        """All six arrays of OFClass (oflow.h:84-86); the last two are only used with usefbcon."""
        self._ck(lib().ofdis_upload_level_fb(self._h, frame, level, _ptr(i0), _ptr(i0x), _ptr(i0y), _ptr(i1), _ptr(i1x),
                                             _ptr(i1y), memkind))

    def upload_pyramids(self, frame: int, pyr):
        for lv in range(self.prm.sc_l, self.prm.sc_f + 1):
            if self.prm.usefbcon:
                self.upload_level_fb(frame, lv, pyr.i0[lv], pyr.i0x[lv], pyr.i0y[lv], pyr.i1[lv], pyr.i1x[lv], pyr.i1y[lv])
            else:
                self.upload_level(frame, lv, pyr.i0[lv], pyr.i0x[lv], pyr.i0y[lv], pyr.i1[lv])

    @property
    def packed_images_frame_floats(self) -> int:
        return lib().ofdis_packed_images_frame_floats(self._h)

    def upload_packed_images(self, f0, f1, packed, memkind=MEM_HOST):
        """I0,I1 only (the leading part of the packed layout); gradients are derived on the device."""
        self._ck(lib().ofdis_upload_packed_images(self._h, f0, f1, _ptr(packed), memkind))

    def get_level(self, frame, level, which) -> np.ndarray:
        """Padded device array `which` (0 I0, 1 I0x, 2 I0y, 3 I1) of one level."""
        h, w = (self.height >> level) + 2 * self.pad, (self.width >> level) + 2 * self.pad
        out = np.empty((h, w) if self.prm.noc == 1 else (h, w, self.prm.noc), np.float32)
        self._ck(lib().ofdis_get_level(self._h, frame, level, which, _ptr(out), MEM_HOST))
        return out

    @property
    def finest_level_frame_floats(self) -> int:
        return lib().ofdis_finest_level_frame_floats(self._h)

    def upload_finest_level(self, f0, f1, packed, memkind=MEM_HOST):
        """[frame][2][h][w][noc] un-padded float images of level sc_l; the rest is derived on the device."""
        self._ck(lib().ofdis_upload_finest_level(self._h, f0, f1, _ptr(packed), memkind))

    def upload_frames_u8(self, f0, f1, frames, width_org, height_org, memkind=MEM_HOST):
        """[frame][2][height_org][width_org][noc] 8-bit pairs; whole pyramid built on the device."""
        self._ck(lib().ofdis_upload_frames_u8(self._h, f0, f1, _ptr(frames), width_org, height_org, memkind))

    def get_flow_fullres(self, f0, f1, dst, width_org, height_org, memkind=MEM_HOST):
        """Flow x 2^sc_l, upsampled to the original frame size and cropped (run_dense.cpp:407-414)."""
        self._ck(lib().ofdis_get_flow_fullres(self._h, f0, f1, _ptr(dst), width_org, height_org, memkind))

    def upload_packed(self, f0, f1, packed, memkind=MEM_HOST):
        self._ck(lib().ofdis_upload_packed(self._h, f0, f1, _ptr(packed), memkind))

    def set_flow(self, frame, level, flow, memkind=MEM_HOST):
        if isinstance(flow, np.ndarray):
            flow = np.ascontiguousarray(flow, np.float32)
        self._ck(lib().ofdis_set_flow(self._h, frame, level, _ptr(flow), memkind))
        if memkind == MEM_HOST:
            self.sync()

    def get_flow(self, frame, level) -> np.ndarray:
        h, w = self.height >> level, self.width >> level
        out = np.empty((h, w, self.prm.nop), np.float32)
        self._ck(lib().ofdis_get_flow(self._h, frame, level, _ptr(out), MEM_HOST))
        return out

    def get_flow_batch(self, f0, f1, dst, memkind=MEM_HOST):
        self._ck(lib().ofdis_get_flow_batch(self._h, f0, f1, _ptr(dst), memkind))

    def get_patches(self, frame, level):
        li = self.level_info(level)
        n_p = li["nopw"] * li["noph"]
        novals = self.prm.noc * self.prm.p_samp_s ** 2
        p = np.empty((n_p, self.prm.nop), np.float32)
        pw = np.empty((n_p, novals), np.float32)
        conv = np.empty(n_p, np.int32)
        cnt = np.empty(n_p, np.int32)
        self._ck(lib().ofdis_get_patches(self._h, frame, level, _ptr(p), _ptr(pw), _ptr(conv), _ptr(cnt)))
        return dict(p=p, pweight=pw, conv=conv, cnt=cnt, **li)

    # -- stage operators ------------------------------------------------------
    def patgrid_optimize(self, level, f0=0, f1=1, init_from_coarser=True):
        self._ck(lib().ofdis_patgrid_optimize(self._h, level, f0, f1, 1 if init_from_coarser else 0))

    def patgrid_aggregate(self, level, f0=0, f1=1):
        self._ck(lib().ofdis_patgrid_aggregate(self._h, level, f0, f1))

    def varref_refine(self, level, f0=0, f1=1, n_inner=None):
        if n_inner is None:
            self._ck(lib().ofdis_varref_refine(self._h, level, f0, f1))
        else:
            self._ck(lib().ofdis_debug_varref_iters(self._h, level, f0, f1, n_inner))

    def run(self, nframes=1, use_initflow=False):
        self._ck(lib().ofdis_run(self._h, nframes, 1 if use_initflow else 0))

    def sync(self):
        self._ck(lib().ofdis_sync(self._h))

    def set_camlr(self, camlr: int):
        self._ck(lib().ofdis_set_camlr(self._h, camlr))

    def set_option(self, name: str, value: int):
        """Launch-geometry options of ofdis_set_option (results are bit-identical under every setting)."""
        self._ck(lib().ofdis_set_option(self._h, name.encode(), int(value)))

    def set_graph_mode(self, on: bool):
        self._ck(lib().ofdis_set_graph_mode(self._h, 1 if on else 0))

    def profile_kernels(self, nframes: int, steps: int = 5):
        """Eager runs with CUDA events around each stage: {class: ms_per_step, launches_per_step}."""
        ms = (ctypes.c_double * 5)()
        n = (ctypes.c_long * 5)()
        self._ck(lib().ofdis_profile_run(self._h, nframes, steps, ms, n))
        names = ("patch", "densify", "vr_setup", "assemble", "sor")
        return {k: {"ms_per_step": ms[i] / steps, "launches_per_step": n[i] / steps} for i, k in enumerate(names)}

    def profile_levels(self, nframes: int, steps: int = 3):
        """Like profile_kernels, split by pyramid level: {level: {class: ms_per_step}}."""
        nlev = self.prm.sc_f - self.prm.sc_l + 1
        ms = (ctypes.c_double * 5)()
        n = (ctypes.c_long * 5)()
        lv = (ctypes.c_double * (5 * nlev))()
        self._ck(lib().ofdis_profile_levels(self._h, nframes, steps, ms, n, lv))
        names = ("patch", "densify", "vr_setup", "assemble", "sor")
        return {self.prm.sc_l + i: {k: lv[i * 5 + j] / steps for j, k in enumerate(names)} for i in range(nlev)}

    @property
    def launch_count(self) -> int:
        return lib().ofdis_launch_count(self._h)

    def debug_get(self, name: str, frame: int, level: int) -> np.ndarray:
        li = self.level_info(level)
        pitch = (li["w"] + 3) // 4 * 4
        C = self.prm.noc
        per = {"mask": 1, "dudv": 2, "rec": 8 if self.prm.nop == 2 else 5}.get(name, C)
        buf = np.empty(pitch * li["h"] * per, np.float32)
        n = lib().ofdis_debug_get(self._h, name.encode(), frame, _ptr(buf), buf.size)
        if n < 0:
            raise OfdisError("debug_get(%s) failed: %d" % (name, n))
        if name in ("dudv", "rec"):
            return buf.reshape(li["h"], pitch, per)[:, :li["w"]]
        return buf.reshape(per, li["h"], pitch)[:, :, :li["w"]]


# ---------------------------------------------------------------------------
# Reference-shaped classes (same names, same argument meaning).
# ---------------------------------------------------------------------------
class OFClass:
    """OFC::OFClass (oflow.h:84-111): all work happens in the constructor; the flow of
    level sc_l is written into `outflow` (numpy, (h, w, nop) float32)."""

    def __init__(self, im_ao, im_ao_dx, im_ao_dy, im_bo, im_bo_dx, im_bo_dy, imgpadding, outflow, initflow, width,
                 height, sc_f, sc_l, max_iter, min_iter, dp_thresh, dr_thresh, res_thresh, p_samp_s, patove, usefbcon,
                 costfct, noc, patnorm, usetvref, tv_alpha, tv_gamma, tv_delta, tv_innerit, tv_solverit, tv_sor,
                 verbosity, nop=2, device=0):
        prm = DisParams(sc_f=sc_f, sc_l=sc_l, max_iter=max_iter, min_iter=min_iter, dp_thresh=dp_thresh,
                        dr_thresh=dr_thresh, res_thresh=res_thresh, p_samp_s=p_samp_s, patove=patove,
                        usefbcon=int(usefbcon), costfct=costfct, noc=noc, patnorm=patnorm, usetvref=int(usetvref),
                        tv_alpha=tv_alpha, tv_gamma=tv_gamma, tv_delta=tv_delta, tv_innerit=tv_innerit,
                        tv_solverit=tv_solverit, tv_sor=tv_sor, verbosity=verbosity, nop=nop)
        ctx = Context(prm, width, height, imgpadding, 1, device)
        try:
            for lv in range(sc_l, sc_f + 1):
                if usefbcon:  # the backward grid's template gradients (oflow.cpp:193-197)
                    ctx.upload_level_fb(0, lv, im_ao[lv], im_ao_dx[lv], im_ao_dy[lv], im_bo[lv], im_bo_dx[lv], im_bo_dy[lv])
                else:
                    ctx.upload_level(0, lv, im_ao[lv], im_ao_dx[lv], im_ao_dy[lv], im_bo[lv])
            if initflow is not None:
                ctx.set_flow(0, sc_f + 1, initflow)
            ctx.run(1, use_initflow=initflow is not None)
            outflow[...] = ctx.get_flow(0, sc_l).reshape(outflow.shape)
        finally:
            ctx.close()


class PatGridClass:
    """OFC::PatGridClass (patchgrid.h:19-44) on one pyramid level of a Context."""

    def __init__(self, ctx: Context, level: int, frame: int = 0):
        self.ctx, self.level, self.frame = ctx, level, frame
        self._i0 = self._i1 = None
        self._from_coarser = False

    def InitializeGrid(self, im_ao, im_ao_dx, im_ao_dy):
        self._i0 = (im_ao, im_ao_dx, im_ao_dy)

    def SetTargetImage(self, im_bo, im_bo_dx=None, im_bo_dy=None):
        self._i1 = im_bo
        self.ctx.upload_level(self.frame, self.level, self._i0[0], self._i0[1], self._i0[2], im_bo)

    def InitializeFromCoarserOF(self, flow_prev):
        self.ctx.set_flow(self.frame, self.level + 1, flow_prev)
        self._from_coarser = True

    def Optimize(self):
        self.ctx.patgrid_optimize(self.level, self.frame, self.frame + 1, self._from_coarser)

    def AggregateFlowDense(self, flowout):
        self.ctx.patgrid_aggregate(self.level, self.frame, self.frame + 1)
        flowout[...] = self.ctx.get_flow(self.frame, self.level).reshape(flowout.shape)

    def GetNoPatches(self):
        li = self.ctx.level_info(self.level)
        return li["nopw"] * li["noph"]

    def GetNopw(self):
        return self.ctx.level_info(self.level)["nopw"]

    def GetNoph(self):
        return self.ctx.level_info(self.level)["noph"]

    def GetRefPatchPos(self, i):
        li = self.ctx.level_info(self.level)
        offw = (li["w"] - (li["nopw"] - 1) * li["steps"]) // 2
        offh = (li["h"] - (li["noph"] - 1) * li["steps"]) // 2
        x, y = divmod(i, li["noph"])
        return np.array([x * li["steps"] + offw, y * li["steps"] + offh], np.float32)

    def GetQuePatchDis(self, i):
        """pt_ref - pt_iter (patchgrid.h:44)."""
        p = self.ctx.get_patches(self.frame, self.level)["p"][i]
        ref = self.GetRefPatchPos(i)
        que = ref.copy()
        que[:len(p)] = ref[:len(p)] + p
        return ref - que


class VarRefClass:
    """OFC::VarRefClass (refine_variational.h:37-39): refines `flowout` in place."""

    def __init__(self, ctx: Context, level: int, flowout: np.ndarray, frame: int = 0):
        ctx.set_flow(frame, level, flowout)
        ctx.varref_refine(level, frame, frame + 1)
        flowout[...] = ctx.get_flow(frame, level).reshape(flowout.shape)
