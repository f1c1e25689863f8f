"""Trajectory iterators over GSD files: which frames to analyse and which origins are active.

Drop-in for reference ``read/_gsd.py`` on top of this package's own GSD reader (``gsdio``): same
function names, arguments and yielded ``(indexes, frame)`` pairs, including the reference's
behaviour on step mismatches and its ``keyframe_max + 1`` origins in linear mode (SURVEY trap T5).
"""

import logging
from pathlib import Path
from typing import Iterator, List, Optional, Tuple

from ..frame import Frame, HoomdFrame
from ..gsdio import GSDTrajectory, open_gsd
from ..StepSize import GenerateStepSeries

logger = logging.getLogger(__name__)

FileIterator = Iterator[Tuple[List[int], Frame]]


def iter_trajectory(trajectory: GSDTrajectory) -> Iterator[HoomdFrame]:
    """Frames of a trajectory, skipping corrupt ones (reference ``_gsd.py:31-52``)."""
    for index in range(len(trajectory)):
        try:
            yield HoomdFrame(trajectory[index])
        except RuntimeError:
            logger.info("Found corrupt frame at index %s. Continuing", index)


def _get_num_steps(trajectory: GSDTrajectory) -> int:
    """Timestep of the last readable frame, retrying up to 10 frames back (reference ``_gsd.py:55-88``)."""
    frame_index = -1
    for _ in range(10):
        try:
            return trajectory[frame_index].step
        except RuntimeError:
            frame_index -= 1
        except IndexError as e:
            logger.error(e)
            raise e
    raise RuntimeError("Cannot read frames from trajectory ", trajectory)


def _gsd_linear_trajectory(
    infile: Path, steps_max: Optional[int] = None, keyframe_interval: int = 20000, keyframe_max: int = 500
):
    """Every frame, one shared growing index list (reference ``_gsd.py:91-116``)."""
    index_list: List[int] = []
    with open_gsd(Path(infile)) as src:
        for frame in iter_trajectory(src):
            if frame.timestep % keyframe_interval == 0 and len(index_list) <= keyframe_max:
                index_list.append(len(index_list))
            if steps_max is not None and frame.timestep > steps_max:
                return
            yield index_list, frame


def _gsd_exponential_trajectory(
    infile: Path,
    steps_max: Optional[int] = None,
    keyframe_interval: int = 20000,
    keyframe_max: int = 500,
    linear_steps: int = 100,
):
    """Frames aligned against ``GenerateStepSeries`` (reference ``_gsd.py:119-189``)."""
    log_level = logger.warning
    with open_gsd(Path(infile)) as src:
        if steps_max is None:
            steps_max = _get_num_steps(src)
        step_iter = GenerateStepSeries(
            steps_max, num_linear=linear_steps, gen_steps=keyframe_interval, max_gen=keyframe_max
        )
        traj_iter = iter_trajectory(src)
        for frame in traj_iter:
            try:
                curr_step = int(next(step_iter))
            except StopIteration:
                return
            if curr_step > frame.timestep:
                # the schedule is ahead: skip trajectory frames
                log_level("Step missing in iterator: iterator %d, trajectory %d", curr_step, frame.timestep)
                log_level = logger.debug
                while frame.timestep < curr_step:
                    try:
                        frame = next(traj_iter)
                    except StopIteration:
                        return
            elif curr_step < frame.timestep:
                # the trajectory is ahead: skip schedule steps
                log_level("Step missing in gsd trajectory: iterator %d, trajectory %d", curr_step, frame.timestep)
                log_level = logger.debug
                while curr_step < frame.timestep:
                    try:
                        curr_step = next(step_iter)
                    except StopIteration:
                        return
            if steps_max is not None and curr_step > steps_max:
                return
            if curr_step == frame.timestep:
                yield step_iter.get_index(), frame


def read_gsd_trajectory(
    infile: Path,
    steps_max: Optional[int] = None,
    linear_steps: Optional[int] = 100,
    keyframe_interval: int = 1_000_000,
    keyframe_max: int = 500,
) -> FileIterator:
    """Reference ``_gsd.py:192-209``."""
    infile = Path(infile)
    if linear_steps is None:
        yield from _gsd_linear_trajectory(infile, steps_max, keyframe_interval, keyframe_max)
    else:
        yield from _gsd_exponential_trajectory(infile, steps_max, keyframe_interval, keyframe_max, linear_steps)
