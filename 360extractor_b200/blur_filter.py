"""The blur accept/reject state machine, replayed over a batch of scores.

In the reference the decision is interleaved with the pixel work, one view at a time
(src/core/processor.py:341-376), and its state (``blur_history``, ``consecutive_blur_skips``,
processor.py:224-225) is carried across views AND frames in iteration order.  The batched GPU path
computes every score first (possibly on several GPUs), so the same decisions are obtained by
replaying the machine over the gathered ``[frame][view]`` scores in that order.
"""
from __future__ import annotations

from collections import deque

HISTORY_LEN = 10            # deque(maxlen=10), processor.py:224
ADAPTIVE_RATIO = 0.6        # score < 0.6 * rolling mean => blurry, processor.py:354
MAX_CONSECUTIVE_SKIPS = 5   # "> 5" consecutive skips forces an accept, processor.py:360


class BlurFilter:
    """Stateful filter; ``accept(score)`` returns True when the view is kept."""

    def __init__(self, enabled=False, smart=False, threshold=100.0):
        self.enabled = bool(enabled)
        self.smart = bool(smart)
        self.threshold = threshold
        self.history = deque(maxlen=HISTORY_LEN)
        self.consecutive_skips = 0
        self.skipped = 0
        self.forced = 0

    @classmethod
    def from_settings(cls, settings):
        """Keys and defaults of src/core/processor.py:219-221."""
        return cls(settings.get("blur_filter_enabled", False),
                   settings.get("smart_blur_enabled", False),
                   settings.get("blur_threshold", 100.0))

    def accept(self, score) -> bool:
        if not self.enabled:                       # processor.py:341 — no score, view always kept
            return True
        blurry = False
        if self.smart:
            if score < self.threshold:             # 1. floor
                blurry = True
            elif len(self.history) > 0:            # 2. adaptive check against the rolling mean
                if score < (sum(self.history) / len(self.history)) * ADAPTIVE_RATIO:
                    blurry = True
            if blurry:                             # 3. safety override
                self.consecutive_skips += 1
                if self.consecutive_skips > MAX_CONSECUTIVE_SKIPS:
                    blurry = False
                    self.consecutive_skips = 0
                    self.forced += 1
            if not blurry:                         # 4. accepted scores feed the history
                self.consecutive_skips = 0
                self.history.append(score)
        else:
            blurry = score < self.threshold        # standard mode, processor.py:369-371
        if blurry:
            self.skipped += 1
        return not blurry

    def replay(self, scores):
        """``scores``: iterable of per-frame iterables in view order -> nested list of keep flags."""
        return [[self.accept(float(s)) for s in row] for row in scores]


def blur_replay(scores, enabled=True, smart=True, threshold=100.0, state=None):
    """The accept / reject machine over a whole ``[F][V]`` score table in (frame, view) order, by the library's C restatement
    (``x360_blur_replay``; the same decisions as ``BlurFilter.replay``, tests/test_host_api.py) -> bool ``[F][V]``.
    ``state``: a ``_capi.BlurState`` to carry on from (and update) instead of a fresh machine."""
    import ctypes

    import numpy as np

    from ._capi import BlurState, check, lib
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    if state is None:
        state = BlurState()
        lib().x360_blur_state_init(state, int(enabled), int(smart), float(threshold))
    keep = np.zeros(scores.size, np.uint8)
    check(lib().x360_blur_replay(ctypes.byref(state), scores.ctypes.data, int(scores.size), keep.ctypes.data))
    return keep.reshape(scores.shape).astype(bool)


__all__ = ["BlurFilter", "blur_replay"]
