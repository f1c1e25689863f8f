"""Step schedules: which timesteps are analysed and which time-origins are active at each.

Host-side drop-in for reference ``StepSize.py`` (same names and semantics: ``generate_steps``,
``exp_sequence``, ``GenerateStepSeries`` with ``get_index()``).  It drives which origins the fused
kernel touches for every frame; it is pure integer bookkeeping and stays in Python.
"""

import heapq
import itertools
from typing import Dict, Iterator, List, Tuple


def exp_sequence(start: int = 0, num_linear: int = 100, initial_step_size: int = 1, base: int = 10) -> Iterator[int]:
    """``start``, then runs of ``num_linear`` equal steps whose size grows by ``base`` after each run
    (reference ``StepSize.py:53-91``).  ``exp_sequence(0, 10)`` -> 0,1,..,10,20,..,100,200,..."""
    yield start
    reached = start
    size = initial_step_size
    while True:
        for multiple in range(num_linear + 1):
            candidate = start + multiple * size
            if candidate > reached:
                yield candidate
        reached = start + num_linear * size
        size *= base


def generate_steps(total_steps: int, num_linear: int = 100, start: int = 0) -> Iterator[int]:
    """The exponential sequence from ``start`` up to, and always ending with, ``total_steps``
    (reference ``StepSize.py:18-50``)."""
    yield from itertools.takewhile(lambda step: step < total_steps, exp_sequence(start, num_linear, 1, 10))
    yield total_steps


class GenerateStepSeries:
    """Merged step series of many time-origins (reference ``StepSize.py:94-171``).

    Origin 0 starts at step 0; a further origin starts at every later step that is a multiple of
    ``gen_steps`` until ``max_gen`` exist.  Iterating yields the distinct steps in increasing order;
    ``get_index()`` returns the origins whose own series contains the step just yielded.
    """

    def __init__(self, total_steps: int, num_linear: int = 100, gen_steps: int = 200000, max_gen: int = 500) -> None:
        self.total_steps = total_steps
        self.num_linear = num_linear
        self.gen_steps = gen_steps
        self.max_gen = max_gen
        self.curr_step = 0
        self._num_generators = 0
        self._pending: List[int] = []  # heap of distinct future steps
        self.values: Dict[int, List[Tuple[int, Iterator[int]]]] = {}
        self._add_generator()

    def _schedule(self, origin: int, series: Iterator[int]) -> None:
        step = next(series, None)
        if step is None:
            return
        waiting = self.values.get(step)
        if waiting is None:
            self.values[step] = [(origin, series)]
            heapq.heappush(self._pending, step)
        else:
            waiting.append((origin, series))

    def _add_generator(self) -> None:
        self._schedule(self._num_generators, generate_steps(self.total_steps, self.num_linear, self.curr_step))
        self._num_generators += 1

    def __iter__(self):
        return self

    def __next__(self) -> int:
        if not self._pending:
            raise StopIteration
        last = self.curr_step
        self.curr_step = heapq.heappop(self._pending)
        if self.curr_step != last:
            self.values.pop(last, None)
        if self.curr_step > 0 and self.curr_step % self.gen_steps == 0 and self._num_generators < self.max_gen:
            self._add_generator()
        for origin, series in tuple(self.values[self.curr_step]):
            self._schedule(origin, series)
        return self.curr_step

    def next(self) -> int:
        return self.__next__()

    def get_index(self) -> List[int]:
        """Origins active at the step most recently returned."""
        return [origin for origin, _ in self.values.get(self.curr_step, ())]
