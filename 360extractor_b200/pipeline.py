"""Per-job driver for the hot path: the caller contract of ``ProcessingWorker.process_video``.

Covers exactly the slices of src/core/processor.py that sit on the path:
  :165-171,199-200,219-221  settings read (resolution, fov, camera_count, pitch_offset, layout_mode,
                            interpolation_mode, blur keys)
  :245-265                  views + one "map" per active camera        -> one libx360 plan
  :327-342                  per frame, per view: remap (+ blur score)  -> one batched GPU call
  :341-376                  blur accept/reject, in (frame, view) order -> ``BlurFilter`` replay
  :228-229,379-381          unsharp mask on the accepted views           -> on the device, inside the batched call
                                                                          (``Extractor(sharpen_strength=...)``)
Decode, sharpening, AI masking, naming and saving stay with the caller (out of scope).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .blur_filter import BlurFilter
from .extractor import Extractor
from .geometry import GeometryProcessor, rotation_stack
from .sharding import gather_sums, gather_sums_by_view, shard_axis, shard_bounds


@dataclass
class ViewResult:
    frame_index: int          # index into the batch passed to ``process``
    camera: str               # view name, used as the camera token of the output file name (processor.py:446)
    image: np.ndarray         # uint8 [d][d][3] BGR, C-contiguous, owned by the caller
    score: Optional[float]    # Laplacian variance, None when the blur filter is off


class ExtractionSession:
    """One job: settings dict + source size -> plan; ``process(frames)`` -> kept views."""

    def __init__(self, settings: dict, src_h: int, src_w: int, device: int = 0, max_batch: int = 0):
        self.settings = settings
        self.out_res = settings.get("resolution", 2048)                 # job.py:40
        self.fov = settings.get("fov", 90)                              # processor.py:166
        camera_count = settings.get("camera_count", 6)                  # processor.py:167
        pitch_offset = settings.get("pitch_offset", 0)                  # processor.py:168
        layout_mode = settings.get("layout_mode", "ring")               # processor.py:169
        if layout_mode == "adaptive":                                   # legacy alias, processor.py:170-171
            layout_mode = "ring"
        self.interpolation_mode = settings.get("interpolation_mode", "linear")
        self.blur = BlurFilter.from_settings(settings)           # host replay (sharded / encoded paths); process() uses the library's copy
        from ._capi import BlurState, lib
        self._blur_state = BlurState()
        lib().x360_blur_state_init(self._blur_state, int(self.blur.enabled), int(self.blur.smart), float(self.blur.threshold))
        self.sharpen_enabled = settings.get("sharpening_enabled", False)   # processor.py:228
        self.sharpen_strength = settings.get("sharpening_strength", 0.5)   # processor.py:229
        self.device = int(device)
        self._max_batch = int(max_batch)
        # flat (non-360) media are passed through untouched by the reference (processor.py:336-338: one "Flat" view =
        # frame.copy()); there is nothing to reproject, so this session refuses them instead of reprojecting silently
        if not settings.get("is_360", True):
            raise ValueError("ExtractionSession reprojects 360 media; settings['is_360'] is False (flat media are copied "
                             "through by the reference, src/core/processor.py:336-338 — no GPU path is needed for that)")
        self.views = GeometryProcessor.generate_views(camera_count, pitch_offset=pitch_offset, layout_mode=layout_mode)
        self.active_cameras = settings.get("active_cameras", None)      # job.py:12
        self.names, R = rotation_stack(self.views, self.active_cameras)
        self.src_h, self.src_w = int(src_h), int(src_w)
        self.extractor = None
        if len(self.names):
            # sharpening_enabled: every chunk is unsharp-masked on the device between the kernel and the download
            # (processor.py:379-381); the reference sharpens only the views it keeps, which are the ones returned
            self.extractor = Extractor(src_h, src_w, self.out_res, self.fov, R, self.interpolation_mode,
                                       device=device, max_batch=max_batch,
                                       sharpen_strength=self.sharpen_strength if self.sharpen_enabled else None)

    def close(self):
        if self.extractor is not None:
            self.extractor.close()
            self.extractor = None
        if getattr(self, "_vs_extractor", None) is not None:
            self._vs_extractor.close()
            self._vs_extractor = None
            self._vs_key = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def extract_batch(self, frames):
        """Raw batch call: ``(views uint8 [F][V][d][d][3], scores float64 [F][V] or None)``."""
        if self.extractor is None:
            F = len(frames)
            return np.empty((F, 0, self.out_res, self.out_res, 3), np.uint8), (np.empty((F, 0)) if self.blur.enabled else None)
        return self.extractor.extract(frames, want_scores=self.blur.enabled)

    def process(self, frames) -> List[ViewResult]:
        """Batch equivalent of the per-frame view loop: returns the kept views in (frame, view) order.

        With the blur filter on, the accept / reject machine runs inside the library's pipeline (``x360_extract_host_filtered``:
        the scores of a chunk come back first, the machine is replayed in (frame, view) order, only the kept views are
        downloaded); its state lives in ``self._blur_state`` across calls, as the reference's does across frames."""
        if self.extractor is None:
            return []
        kept_views, keep, scores = self.extractor.extract_filtered(frames, self._blur_state)
        self.blur.skipped, self.blur.forced = int(self._blur_state.skipped), int(self._blur_state.forced)
        kept: List[ViewResult] = []
        i = 0
        for f in range(keep.shape[0]):
            for v, name in enumerate(self.names):
                if keep[f, v]:
                    kept.append(ViewResult(f, name, kept_views[i], float(scores[f, v]) if scores is not None else None))
                    i += 1
        return kept

    def process_encoded(self, frames, quality=None):
        """``process`` with the save step's encode on the device: the kept views come back as the ``.jpg`` files
        ``FileManager.save_image`` would write (src/core/processor.py:121-123,428-431; quality = settings 'quality', default 95),
        byte-identical to cv2's, and only compressed bytes cross PCIe.  Returns ``[(frame_index, camera, jpeg_bytes, score)]``."""
        if quality is None:
            quality = self.settings.get("quality", 95)
        if self.extractor is None:
            return []
        files, scores = self.extractor.extract_jpeg(frames, quality=quality, want_scores=self.blur.enabled)
        kept = []
        for f in range(len(files)):
            for v, name in enumerate(self.names):
                score = float(scores[f, v]) if scores is not None else None
                if self.blur.accept(score):
                    kept.append((f, name, files[f][v], score))
        return kept

    def process_sharded(self, frames, rank: int, world_size: int, group=None):
        """Frame-sharded variant: this rank reprojects frames ``shard_bounds(F, world, rank)``, the
        score sums are all-gathered and every rank replays the blur machine over the full table.

        Returns ``(lo, local_views, keep)`` with ``keep`` a bool ``[F][V]`` table valid on every rank.
        Jobs with fewer frames than ranks (single images) are sharded by view instead: see
        ``process_view_sharded``."""
        F = len(frames)
        lo, hi = shard_bounds(F, world_size, rank)
        local = np.ascontiguousarray(frames[lo:hi])
        V = len(self.names)
        if self.extractor is None or not self.blur.enabled:
            views = self.extract_batch(local)[0] if hi > lo else np.empty((0, V, self.out_res, self.out_res, 3), np.uint8)
            return lo, views, np.ones((F, V), bool)
        if hi > lo:
            views, s, s2 = self.extractor.extract_sums(local)
        else:
            views = np.empty((0, V, self.out_res, self.out_res, 3), np.uint8)
            s = s2 = np.zeros((0, V), np.int64)
        gs, gs2 = gather_sums(s, s2, F, group=group)
        from .extractor import scores_from_sums
        scores = scores_from_sums(gs, gs2, self.extractor.pixels_per_view)
        keep = np.array(self.blur.replay(scores), dtype=bool).reshape(F, V)
        return lo, views, keep

    def _view_shard_extractor(self, v_lo: int, v_hi: int):
        """A plan over this rank's block of the emitted views (same frame size, same settings)."""
        key = (v_lo, v_hi)
        if getattr(self, "_vs_key", None) != key:
            if getattr(self, "_vs_extractor", None) is not None:
                self._vs_extractor.close()
            _, R = rotation_stack(self.views, self.active_cameras)
            self._vs_extractor = Extractor(self.src_h, self.src_w, self.out_res, self.fov, R[v_lo:v_hi], self.interpolation_mode,
                                           device=self.device, max_batch=self._max_batch,
                                           sharpen_strength=self.sharpen_strength if self.sharpen_enabled else None)
            self._vs_key = key
        return self._vs_extractor

    def process_view_sharded(self, frames, rank: int, world_size: int, group=None):
        """View-sharded variant for jobs with fewer frames than GPUs (a single 360 image: the loop of
        src/core/processor.py:327-330 is what gets split): this rank reprojects EVERY frame for the views
        ``shard_bounds(V, world, rank)`` of the emitted view list; sums are all-gathered along the view axis and
        the blur machine is replayed in (frame, view) order on every rank.

        Returns ``(v_lo, local_views [F][V_local][d][d][3], keep [F][V])``."""
        F, V = len(frames), len(self.names)
        v_lo, v_hi = shard_bounds(V, world_size, rank)
        frames = np.ascontiguousarray(frames)
        if v_hi == v_lo:
            views = np.empty((F, 0, self.out_res, self.out_res, 3), np.uint8)
            s = s2 = np.zeros((F, 0), np.int64)
        else:
            ex = self._view_shard_extractor(v_lo, v_hi)
            if self.blur.enabled:
                views, s, s2 = ex.extract_sums(frames)
            else:
                views, s, s2 = ex.extract(frames)[0], None, None
        if not self.blur.enabled:
            return v_lo, views, np.ones((F, V), bool)
        gs, gs2 = gather_sums_by_view(s, s2, V, group=group)
        from .extractor import scores_from_sums
        scores = scores_from_sums(gs, gs2, self.out_res * self.out_res)
        keep = np.array(self.blur.replay(scores), dtype=bool).reshape(F, V)
        return v_lo, views, keep

    def process_distributed(self, frames, rank: int, world_size: int, group=None):
        """Pick the sharding axis: frames when there are at least as many as ranks, else views.
        Returns ``(axis, lo, local_views, keep)``."""
        axis = shard_axis(len(frames), world_size)
        if axis == "frame":
            lo, views, keep = self.process_sharded(frames, rank, world_size, group=group)
        else:
            lo, views, keep = self.process_view_sharded(frames, rank, world_size, group=group)
        return axis, lo, views, keep


__all__ = ["ExtractionSession", "ViewResult"]
