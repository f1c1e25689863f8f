"""Restatement of the reference step schedule (``src/sdanalysis/StepSize.py:18-171``).
TEST INFRASTRUCTURE ONLY.  Pinned against sequences produced by importing the reference's own
``StepSize.py`` (stdlib-only, importable here): ``tests/golden/stepsize_golden.json``.
"""

import heapq


def exp_sequence(start=0, num_linear=100, initial_step_size=1, base=10):
    """StepSize.py:53-91: ``num_linear`` equal steps, then the step grows by ``base``."""
    size = initial_step_size
    current = start
    yield current
    while True:
        for i in range(num_linear + 1):
            value = start + i * size
            if value > current:
                yield value
        current = start + num_linear * size
        size *= base


def generate_steps(total_steps, num_linear=100, start=0):
    """StepSize.py:18-50: the exponential sequence below ``total_steps``, then ``total_steps``."""
    for value in exp_sequence(start, num_linear, 1, 10):
        if not value < total_steps:
            break
        yield value
    yield total_steps


class GenerateStepSeries:
    """StepSize.py:94-171: merge of one ``generate_steps`` series per time-origin.

    A new origin starts at every step that is a positive multiple of ``gen_steps`` while fewer than
    ``max_gen`` exist; ``get_index()`` lists, in creation order of their arrival at this step, the
    origins whose own series contains the current step.
    """

    def __init__(self, total_steps, num_linear=100, gen_steps=200000, max_gen=500):
        self.total_steps = total_steps
        self.num_linear = num_linear
        self.gen_steps = gen_steps
        self.max_gen = max_gen
        self.curr_step = 0
        self._count = 0
        self._at = {}  # step -> [(origin index, iterator)]
        self._heap = []
        self._spawn()

    def _advance(self, entry):
        try:
            step = next(entry[1])
        except StopIteration:
            return
        if step in self._at:
            self._at[step].append(entry)
        else:
            self._at[step] = [entry]
            heapq.heappush(self._heap, step)

    def _spawn(self):
        self._advance((self._count, generate_steps(self.total_steps, self.num_linear, self.curr_step)))
        self._count += 1

    def __iter__(self):
        return self

    def __next__(self):
        previous = self.curr_step
        if not self._heap:
            raise StopIteration
        self.curr_step = heapq.heappop(self._heap)
        if self.curr_step != previous:
            del self._at[previous]
        if self.curr_step % self.gen_steps == 0 and self.curr_step > 0 and self._count < self.max_gen:
            self._spawn()
        for entry in list(self._at[self.curr_step]):
            self._advance(entry)
        return self.curr_step

    next = __next__

    def get_index(self):
        return [index for index, _ in self._at.get(self.curr_step, [])]
