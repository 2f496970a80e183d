"""``Extractor``: one libx360 plan = one job's geometry, and the batched hot-path call.

Replaces the reference's per-job map build (src/core/processor.py:245-265) and the body of its
per-view loop (src/core/processor.py:327-342: ``cv2.remap`` + ``calculate_blur_score``) with one
call per frame batch.  Every pixel is computed by the CUDA library; nothing here has a CPU path.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _capi
from ._capi import INTERP_LANCZOS4, INTERP_LINEAR, Params, check, lib, ptr


def interp_code(interpolation_mode) -> int:
    """'lanczos' -> INTER_LANCZOS4, anything else -> INTER_LINEAR (src/core/processor.py:199-200)."""
    if isinstance(interpolation_mode, (int, np.integer)):
        return int(interpolation_mode)
    return INTERP_LANCZOS4 if interpolation_mode == "lanczos" else INTERP_LINEAR


def scores_from_sums(sum_lap, sum_lap2, n_pixels: int) -> np.ndarray:
    """Population variance of the Laplacian (ndarray.var, ddof=0; src/utils/image_utils.py:27) from
    the exact integer sums the kernel returns: s2/N - (s/N)^2 in fp64."""
    s = np.asarray(sum_lap, dtype=np.float64)
    s2 = np.asarray(sum_lap2, dtype=np.float64)
    mean = s / float(n_pixels)
    return s2 / float(n_pixels) - mean * mean


def plan_preview(src_h, src_w, out_res, fov_deg, R, interpolation_mode="linear", *, out_h=None, flags=0):
    """Host-only view of how the tiled bilinear path organises a job (no device needed): the view units
    that share one map evaluation (views differing only by a yaw), each view's map_x shift inside its
    unit, and the shared-memory staging capacity chosen from the tiles' source footprints."""
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 9)
    p = Params()
    p.struct_size = ctypes.sizeof(Params)
    p.src_w, p.src_h = int(src_w), int(src_h)
    p.out_w, p.out_h = int(out_res), int(out_h if out_h is not None else out_res)
    p.fov_deg = float(fov_deg)
    p.interp = interp_code(interpolation_mode)
    p.n_views = int(R.shape[0])
    p.R = R.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    p.flags = int(flags)
    n_units = ctypes.c_int32()
    rows, bw, th = (ctypes.c_int32 * 2)(), (ctypes.c_int32 * 2)(), (ctypes.c_int32 * 2)()
    unit = np.zeros(p.n_views, np.int32)
    shift = np.zeros(p.n_views, np.float64)
    check(lib().x360_plan_preview(ctypes.byref(p), ctypes.byref(n_units), unit.ctypes.data, shift.ctypes.data,
                                  rows, bw, th))
    group = np.zeros(p.n_views, np.int32)
    th_view = np.zeros((p.n_views, 2), np.int32)
    check(lib().x360_plan_preview_views(ctypes.byref(p), group.ctypes.data, th_view.ctypes.data))
    return {"n_units": n_units.value, "unit_of_view": unit.tolist(), "shift_px": shift.tolist(),
            "rows_max": rows[0], "bw_max": bw[0], "tile_h": th[0],
            "with_scores": {"rows_max": rows[1], "bw_max": bw[1], "tile_h": th[1]},
            # groups: yaw-only views / the others, each with its own kernel variant and tile height
            "group_of_view": group.tolist(), "tile_h_of_view": th_view[:, 0].tolist(), "tile_h_of_view_fused_score": th_view[:, 1].tolist()}


def column_only_map_x(R) -> bool:
    """Host-only: True if ``map_x`` of every view in ``R`` ([V][3][3]) does not depend on the output row (yaw-only
    rotations) — the plans that take the kernel's one-row ``map_x`` variant."""
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 9)
    return bool(lib().x360_column_only_map_x(R.ctypes.data, int(R.shape[0])))


class Extractor:
    """Plan wrapper.  ``R`` is fp64 ``[V][3][3]`` from ``GeometryProcessor.get_rotation_matrix``."""

    def __init__(self, src_h, src_w, out_res, fov_deg, R, interpolation_mode="linear", *,
                 out_h=None, device=0, max_batch=0, path=_capi.PATH_AUTO, sharpen_strength=None, flags=0):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, 9)
        self.src_h, self.src_w = int(src_h), int(src_w)
        self.out_w = int(out_res)
        self.out_h = int(out_h) if out_h is not None else int(out_res)
        self.n_views = int(R.shape[0])
        self.fov_deg = float(fov_deg)
        self.interp = interp_code(interpolation_mode)
        self.device = int(device)
        self._R = R
        p = Params()
        p.struct_size = ctypes.sizeof(Params)
        p.src_w, p.src_h = self.src_w, self.src_h
        p.out_w, p.out_h = self.out_w, self.out_h
        p.fov_deg = self.fov_deg
        p.interp = self.interp
        p.n_views = self.n_views
        p.R = R.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        p.device = self.device
        p.max_batch = int(max_batch)
        p.path = int(path)
        p.flags = int(flags)              # FLAG_* overrides of the plan's own choices (include/x360.h); 0 = let the plan choose
        self._plan = ctypes.c_void_p()
        check(lib().x360_plan_create(ctypes.byref(p), ctypes.byref(self._plan)))
        self.sharpen_strength = None
        if sharpen_strength is not None:
            self.set_sharpen(sharpen_strength)

    def set_sharpen(self, strength):
        """``strength`` (float) makes ``extract`` / ``extract_sums`` return unsharp-masked views
        (src/core/processor.py:379-381, applied on the device before the download); ``None`` turns it off.
        Scores are always those of the plain views, as in the reference (it filters, then sharpens)."""
        check(lib().x360_plan_set_sharpen(self._plan, 0 if strength is None else 1, 0.0 if strength is None else float(strength)))
        self.sharpen_strength = None if strength is None else float(strength)

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_plan", None):
            lib().x360_plan_destroy(self._plan)
            self._plan = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- shapes --------------------------------------------------------------------------------
    @property
    def frame_shape(self):
        return (self.src_h, self.src_w, 3)

    def views_shape(self, n_frames):
        return (n_frames, self.n_views, self.out_h, self.out_w, 3)

    @property
    def pixels_per_view(self):
        return self.out_h * self.out_w

    def info(self):
        vals = [ctypes.c_int32() for _ in range(4)]
        check(lib().x360_plan_info(self._plan, *[ctypes.byref(v) for v in vals]))
        return {"path": vals[0].value, "grid": vals[1].value, "block": vals[2].value, "smem_bytes": vals[3].value}

    def tile_modes(self):
        """(tile, view) pairs of the last launch per gather mode (synchronises): staged in a regular TMA box / in a wide
        box / across the seam, or gathered directly (aligned words / bytes)."""
        out = (ctypes.c_uint32 * 5)()
        check(lib().x360_plan_tile_modes(self._plan, out))
        return {"staged": int(out[0]), "direct_aligned": int(out[1]), "direct_bytewise": int(out[2]),
                "staged_wide": int(out[3]), "staged_seam": int(out[4])}

    def upload_rows(self):
        """``(first_row, n_rows)``: the band of source rows the host entries upload per frame (the rows the views can read;
        the whole frame for plans that see a pole, and with ``FLAG_FULL_UPLOAD``)."""
        r0, n = ctypes.c_int32(), ctypes.c_int32()
        check(lib().x360_plan_upload_rows(self._plan, ctypes.byref(r0), ctypes.byref(n)))
        return int(r0.value), int(n.value)

    def transposed_blocks(self):
        """32 x 32 output blocks (over all views) the plan gathers with the lanes running down the rows (pole-face quadrants
        where the source row changes faster along x than along y); a layout choice, the pixels are the same."""
        n = ctypes.c_int64()
        check(lib().x360_plan_transposed_blocks(self._plan, ctypes.byref(n)))
        return int(n.value)

    # -- the hot path --------------------------------------------------------------------------
    def _host_frames(self, frames):
        frames = np.asarray(frames)
        if frames.dtype != np.uint8:
            raise TypeError(f"frames must be uint8 (got {frames.dtype}); the reference's frames are always 8-bit BGR")
        frames = np.ascontiguousarray(frames)
        if frames.ndim == 3:
            frames = frames[None]
        if frames.ndim != 4 or frames.shape[1:] != self.frame_shape:
            raise ValueError(f"frames shape {frames.shape} does not match plan {self.frame_shape}")
        return frames, frames.shape[0]

    def _host_batch(self, frames, out):
        """Validate a host frame batch and its output buffer (the C side trusts the sizes it is given)."""
        frames, F = self._host_frames(frames)
        if out is None:
            out = np.empty(self.views_shape(F), np.uint8)
        elif not isinstance(out, np.ndarray) or out.shape != self.views_shape(F) or out.dtype != np.uint8 or not out.flags.c_contiguous \
                or not out.flags.writeable:
            raise ValueError("out must be a writable C-contiguous uint8 array of shape %r" % (self.views_shape(F),))
        return frames, out, F

    def extract(self, frames, want_scores=False, out=None):
        """frames uint8 ``[F][H][W][3]`` (numpy / pinned numpy) -> ``(views uint8 [F][V][d][d][3], scores)``.

        ``scores`` is float64 ``[F][V]`` (or ``None``).  Host buffers: upload, kernels and download
        are pipelined in the library on side streams."""
        frames, out, F = self._host_batch(frames, out)
        s = s2 = None
        if want_scores:
            s = np.zeros((F, self.n_views), np.int64)
            s2 = np.zeros((F, self.n_views), np.int64)
        check(lib().x360_extract_host(self._plan, ptr(frames), F, ptr(out), ptr(s), ptr(s2)))
        scores = scores_from_sums(s, s2, self.pixels_per_view) if want_scores else None
        return out, scores

    def extract_sums(self, frames, out=None):
        """Like ``extract`` but returns the raw int64 ``(sum L, sum L^2)`` arrays."""
        frames, out, F = self._host_batch(frames, out)
        s = np.zeros((F, self.n_views), np.int64)
        s2 = np.zeros((F, self.n_views), np.int64)
        check(lib().x360_extract_host(self._plan, ptr(frames), F, ptr(out), ptr(s), ptr(s2)))
        return out, s, s2

    def extract_filtered(self, frames, state, out=None):
        """The pipelined host path that downloads only the views the blur machine keeps (``state``: a ``_capi.BlurState``
        carried across calls, as the reference carries ``blur_history`` / ``consecutive_blur_skips`` across frames,
        src/core/processor.py:224-225,341-376).  Returns ``(kept_views uint8 [K][d][d][3], keep bool [F][V], scores float64
        [F][V] or None)``; rejected views never cross PCIe."""
        frames, F = self._host_frames(frames)
        n = F * self.n_views
        if out is None:
            out = np.empty((n, self.out_h, self.out_w, 3), np.uint8)
        elif out.dtype != np.uint8 or not out.flags.c_contiguous or out.shape[1:] != (self.out_h, self.out_w, 3):
            raise ValueError("out must be a C-contiguous uint8 array [capacity][d][d][3]")
        keep = np.zeros(n, np.uint8)
        scores = np.zeros(n, np.float64)
        k = ctypes.c_int64(0)
        check(lib().x360_extract_host_filtered(self._plan, ptr(frames), F, ctypes.byref(state), ptr(out), int(out.shape[0]), ptr(keep), ptr(scores),
                                               ctypes.byref(k)))
        return out[: k.value], keep.reshape(F, self.n_views).astype(bool), (scores.reshape(F, self.n_views) if state.enabled else None)

    def jpeg_capacity(self, n_frames):
        """Bytes that always hold the files of ``n_frames`` frames (the raw size of the views plus headers)."""
        return n_frames * self.n_views * (self.out_h * self.out_w * 3 + 1024)

    def extract_jpeg_into(self, frames, out, offsets, quality=95, sum_lap=None, sum_lap2=None):
        """Low-level form of ``extract_jpeg``: the files go back to back into ``out`` (uint8, 1-D, pinned for full speed) and
        their byte offsets into ``offsets`` (int64 ``[F * V + 1]``); returns the total number of bytes."""
        frames, F = self._host_frames(frames)
        n = F * self.n_views
        if out.dtype != np.uint8 or out.ndim != 1 or not out.flags.c_contiguous or offsets.dtype != np.int64 or offsets.size < n + 1:
            raise ValueError("out must be a 1-D uint8 array and offsets an int64 array of at least F * V + 1 entries")
        check(lib().x360_extract_host_jpeg(self._plan, ptr(frames), F, int(quality), ptr(out), int(out.size), ptr(offsets), ptr(sum_lap), ptr(sum_lap2)))
        return int(offsets[n])

    def extract_jpeg(self, frames, quality=95, want_scores=False):
        """Like ``extract`` but the views come back as the ``.jpg`` files the reference's save step would write
        (``cv2.imwrite(path, view, [cv2.IMWRITE_JPEG_QUALITY, quality])``, src/core/processor.py:121-123,
        src/utils/file_manager.py:21-32) — encoded on the device, so only compressed bytes cross PCIe.

        Returns ``(files, scores)``: ``files[f][v]`` is a ``bytes`` object, byte-identical to cv2's encode of the same view."""
        frames, F = self._host_frames(frames)
        n = F * self.n_views
        out = np.empty(self.jpeg_capacity(F), np.uint8)
        off = np.zeros(n + 1, np.int64)
        s = s2 = None
        if want_scores:
            s = np.zeros((F, self.n_views), np.int64)
            s2 = np.zeros((F, self.n_views), np.int64)
        self.extract_jpeg_into(frames, out, off, quality, s, s2)
        files = [[out[off[f * self.n_views + v]:off[f * self.n_views + v + 1]].tobytes() for v in range(self.n_views)] for f in range(F)]
        return files, (scores_from_sums(s, s2, self.pixels_per_view) if want_scores else None)

    def extract_device(self, frames, out, sum_lap=None, sum_lap2=None, stream=None):
        """Device-resident batch (torch CUDA tensors or raw device addresses); asynchronous on ``stream``.

        frames ``[F][H][W][3]`` uint8, out ``[F][V][d][d][3]`` uint8, sums int64 ``[F][V]`` or None."""
        F = int(frames.shape[0])
        for name, t, shape in (("frames", frames, (F,) + self.frame_shape), ("out", out, self.views_shape(F)),
                               ("sum_lap", sum_lap, (F, self.n_views)), ("sum_lap2", sum_lap2, (F, self.n_views))):
            if t is None:
                continue
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
            contiguous = t.is_contiguous() if hasattr(t, "is_contiguous") else t.flags.c_contiguous
            if not contiguous:
                raise ValueError(f"{name} must be C-contiguous (the C-ABI takes dense [F][H][W][3] / [F][V][d][d][3] buffers)")
        st = None
        if stream is not None:
            st = getattr(stream, "cuda_stream", stream)
        check(lib().x360_extract_device(self._plan, ptr(frames), F, ptr(out), ptr(sum_lap), ptr(sum_lap2), st))

    # -- materialised maps (compatibility entry) -----------------------------------------------
    def make_maps(self, view: int):
        mx = np.empty((self.out_h, self.out_w), np.float32)
        my = np.empty((self.out_h, self.out_w), np.float32)
        check(lib().x360_make_maps(self._plan, int(view), ptr(mx), ptr(my)))
        return mx, my


def remap(frame, map_x, map_y, interpolation=INTERP_LINEAR, device=0):
    """GPU ``cv2.remap(frame, map_x, map_y, interp, borderMode=cv2.BORDER_WRAP)`` for caller-held maps
    (src/core/processor.py:334).  ``interpolation``: 'linear' / 'lanczos' or an ``INTERP_*`` code."""
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    if frame.ndim != 3 or frame.shape[2] != 3:
        raise ValueError("frame must be uint8 [H][W][3] (BGR)")
    mx = np.ascontiguousarray(map_x, dtype=np.float32)
    my = np.ascontiguousarray(map_y, dtype=np.float32)
    if mx.shape != my.shape or mx.ndim != 2:
        raise ValueError("map_x / map_y must be 2-D float32 arrays of equal shape")
    dst = np.empty(mx.shape + (3,), np.uint8)
    check(lib().x360_remap_maps(int(device), ptr(frame), frame.shape[0], frame.shape[1], ptr(mx), ptr(my),
                                mx.shape[0], mx.shape[1], interp_code(interpolation), ptr(dst)))
    return dst


class PinnedBuffer:
    """A page-locked host array from ``x360_host_alloc`` (for the pipelined host path)."""

    def __init__(self, shape, dtype=np.uint8):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._ptr = ctypes.c_void_p()
        check(lib().x360_host_alloc(max(nbytes, 1), ctypes.byref(self._ptr)))
        buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._ptr:
            self.array = None
            lib().x360_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
