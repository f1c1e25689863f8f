"""LAMMPS text trajectories (reference ``read/_lammps.py:28-95``): one time-origin, every frame.

Host-side only: a ``.lammpstrj`` file carries positions but no orientations, so the frames reach the
hot path with identity quaternions (reference ``frame.py:104-108``) and the rotational quantities
stay at their zero-motion values.
"""

import logging
from pathlib import Path
from typing import Iterator, List, Optional, Tuple

import numpy as np

from ..frame import LammpsFrame

logger = logging.getLogger(__name__)


def parse_lammpstrj(filename: Path) -> Iterator[LammpsFrame]:
    """Frames of a ``dump custom`` file.  Atom rows are placed by their (1-based) ``id`` column, all
    columns are kept as float32 arrays like the reference; a malformed header raises ``RuntimeError``."""
    logger.debug("Parse file: %s", filename)
    with open(filename) as src:
        while True:
            line = src.readline()
            if not line:
                return  # end of file
            try:
                if line != "ITEM: TIMESTEP\n":
                    raise ValueError(line)
                timestep = int(src.readline().strip())
                if src.readline() != "ITEM: NUMBER OF ATOMS\n":
                    raise ValueError("expected ITEM: NUMBER OF ATOMS")
                num_atoms = int(src.readline().strip())
                if "ITEM: BOX BOUNDS" not in src.readline():
                    raise ValueError("expected ITEM: BOX BOUNDS")
                bounds = [src.readline().split() for _ in range(3)]
                box = np.array([float(hi) - float(lo) for lo, hi, *_ in bounds])
                header = src.readline()
                if "ITEM: ATOMS" not in header:
                    raise ValueError("expected ITEM: ATOMS")
                headings = header.strip().split(" ")[2:]
                id_col = headings.index("id")
                rows = np.array([src.readline().split() for _ in range(num_atoms)], dtype=np.float64)
                if rows.shape != (num_atoms, len(headings)):
                    raise ValueError(f"expected {num_atoms} rows of {len(headings)} columns")
            except (ValueError, IndexError) as exc:
                raise RuntimeError("Unable to parse trajectory. Found error", exc) from exc
            slots = rows[:, id_col].astype(np.int64) - 1  # lammps ids start at 1
            frame = {}
            for col, field in enumerate(headings):
                values = np.empty(num_atoms, dtype=np.float32)
                values[slots] = rows[:, col]
                frame[field] = values
            frame["box"] = box
            frame["timestep"] = timestep
            yield LammpsFrame(frame)


def read_lammps_trajectory(infile: Path, steps_max: Optional[int] = None) -> Iterator[Tuple[List[int], LammpsFrame]]:
    """``(indexes, frame)`` pairs with the single time-origin 0 (reference ``_lammps.py:28-41``)."""
    indexes = [0]
    for frame in parse_lammpstrj(Path(infile)):
        if steps_max is not None and frame.timestep > steps_max:
            return
        yield indexes, frame
