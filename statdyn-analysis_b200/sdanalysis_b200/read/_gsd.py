"""Which frames of a GSD trajectory are analysed, and which time-origins each of them advances.

Host-side stand-in for reference ``read/_gsd.py`` on top of this package's own GSD reader
(``gsdio``): the public names (``read_gsd_trajectory``, ``iter_trajectory``) and the
``(origin indexes, frame)`` pairs they produce follow the reference, including what it does when the
trajectory and the step schedule disagree and the ``keyframe_max + 1`` origins of its linear mode
(SURVEY trap T5).
"""

import logging
import os
from pathlib import Path
from typing import Iterator, List, Optional, Tuple

from ..frame import Frame, HoomdFrame
from ..gsdio import GSDTrajectory, open_gsd
from ..StepSize import GenerateStepSeries

logger = logging.getLogger(__name__)

FileIterator = Iterator[Tuple[List[int], Frame]]
_LAST_FRAME_ATTEMPTS = 10


def iter_trajectory(trajectory: GSDTrajectory) -> Iterator[HoomdFrame]:
    """Readable frames in file order; a frame that fails to decode is logged and left out
    (reference ``_gsd.py:31-52``).  Frames are views of the file mapping; the engine stages them through its
    pinned ring.  ``SDB_PIN_TRAJECTORY=1`` page-locks the whole mapping instead (uploads straight from the page
    cache): that pays when frames are read more than once -- registering costs ~0.1 s per GB (measured), more
    than the staging copy of a single pass."""
    if os.environ.get("SDB_PIN_TRAJECTORY", "0") == "1":
        trajectory.pin()
    for position in range(len(trajectory)):
        try:
            snapshot = trajectory[position]
        except RuntimeError:
            logger.info("Found corrupt frame at index %s. Continuing", position)
            continue
        yield HoomdFrame(snapshot)


def _get_num_steps(trajectory: GSDTrajectory) -> int:
    """Timestep of the last frame; when the tail of the file is damaged, of the last one of the final
    ten frames that can be read (reference ``_gsd.py:55-88``)."""
    for back in range(1, _LAST_FRAME_ATTEMPTS + 1):
        try:
            return trajectory[-back].step
        except RuntimeError:
            continue
        except IndexError as exc:
            logger.error(exc)
            raise
    raise RuntimeError("Cannot read frames from trajectory ", trajectory)


def _gsd_linear_trajectory(
    infile: Path, steps_max: Optional[int] = None, keyframe_interval: int = 20000, keyframe_max: int = 500
) -> FileIterator:
    """Dense mode (reference ``_gsd.py:91-116``): every frame advances every origin opened so far.  An
    origin opens at each multiple of ``keyframe_interval`` until ``keyframe_max + 1`` exist; the list
    handed out is one object that grows in place, as in the reference."""
    origins: List[int] = []
    with open_gsd(Path(infile)) as trajectory:
        for frame in iter_trajectory(trajectory):
            on_interval = frame.timestep % keyframe_interval == 0
            if on_interval and len(origins) < keyframe_max + 1:
                origins.append(len(origins))
            if steps_max is not None and frame.timestep > steps_max:
                break
            yield origins, frame


class _MismatchLog:
    """The first disagreement between schedule and file is a warning, later ones debug messages."""

    def __init__(self):
        self._emit = logger.warning

    def __call__(self, what, scheduled, found):
        self._emit("Step missing in %s: iterator %d, trajectory %d", what, scheduled, found)
        self._emit = logger.debug


def _gsd_exponential_trajectory(
    infile: Path,
    steps_max: Optional[int] = None,
    keyframe_interval: int = 20000,
    keyframe_max: int = 500,
    linear_steps: int = 100,
) -> FileIterator:
    """Log-spaced mode (reference ``_gsd.py:119-189``): the frames are matched against the merged step
    series of all origins, and a frame is handed out together with the origins scheduled at its
    timestep.

    Each round draws one frame and one scheduled step.  If they differ, the side that is behind is run
    forward until it has caught up (frames that were never scheduled, or scheduled steps that were never
    written, are dropped); the pair is used only if the two then agree, otherwise the round ends
    empty-handed and the next one draws afresh from both -- the reference's behaviour when both sides have
    gaps."""
    report = _MismatchLog()
    with open_gsd(Path(infile)) as trajectory:
        last_step = _get_num_steps(trajectory) if steps_max is None else steps_max
        schedule = GenerateStepSeries(
            last_step, num_linear=linear_steps, gen_steps=keyframe_interval, max_gen=keyframe_max
        )
        frames = iter_trajectory(trajectory)
        sentinel = object()
        while True:
            frame = next(frames, sentinel)
            scheduled = next(schedule, sentinel)
            if frame is sentinel or scheduled is sentinel:
                return
            scheduled = int(scheduled)
            if frame.timestep < scheduled:
                report("iterator", scheduled, frame.timestep)
                while frame.timestep < scheduled:
                    frame = next(frames, sentinel)
                    if frame is sentinel:
                        return
            elif frame.timestep > scheduled:
                report("gsd trajectory", scheduled, frame.timestep)
                while scheduled < frame.timestep:
                    scheduled = next(schedule, sentinel)
                    if scheduled is sentinel:
                        return
            if scheduled > last_step:
                return
            if scheduled == frame.timestep:
                yield schedule.get_index(), frame


def read_gsd_trajectory(
    infile: Path,
    steps_max: Optional[int] = None,
    linear_steps: Optional[int] = 100,
    keyframe_interval: int = 1_000_000,
    keyframe_max: int = 500,
) -> FileIterator:
    """``(origin indexes, frame)`` for every analysed frame of ``infile``: dense mode when
    ``linear_steps`` is None, log-spaced otherwise (reference ``_gsd.py:192-209``)."""
    if linear_steps is None:
        return _gsd_linear_trajectory(Path(infile), steps_max, keyframe_interval, keyframe_max)
    return _gsd_exponential_trajectory(Path(infile), steps_max, keyframe_interval, keyframe_max, linear_steps)
