"""The driver loop: read a trajectory and compute the dynamics of every active time-origin.

Drop-in for reference ``read/_read.py``: ``process_file`` keeps the signature, the row schema (15
``compute_all`` quantities + ``keyframe``, ``temperature``, ``pressure``) and the per-row error
handling, but instead of one ``(Dynamics, Relaxations)`` pair of Python objects per origin
(``_read.py:117,135-141``) all origins live in one :class:`~sdanalysis_b200.engine.DynamicsEngine`
and each frame costs one fused kernel pass; result rows come back asynchronously.
"""

import datetime
import logging
from collections import deque
from pathlib import Path
from typing import Any, Dict, Iterator, List, Optional

import numpy as np
import pandas

from ..engine import DynamicsEngine
from ..frame import Frame, HoomdFrame
from ..gsdio import open_gsd
from ..molecules import Trimer
from ..util import get_filename_vars
from ._gsd import FileIterator, read_gsd_trajectory
from ._lammps import parse_lammpstrj, read_lammps_trajectory

logger = logging.getLogger(__name__)


# Open parquet writers of this process, by target path: a parquet file is appended to one row group at a time
# through a writer that stays open until the tables of the run are closed (`close_tables`).
_parquet_writers: Dict[str, Any] = {}


def _table_target(filename: Path, group: str):
    """(kind, path) of the file a table goes to: HDF5 like the reference when PyTables is present
    (``_read.py:57-61``: one file, one group per table); otherwise parquet / CSV -- the ``dynamics`` table goes to
    ``filename`` itself, any other table to ``<stem>.<group><suffix>`` next to it."""
    filename = Path(filename)
    if filename.suffix in (".h5", ".hdf5", ".hdf"):
        try:
            import tables  # noqa: F401

            return "hdf", filename
        except ImportError:
            logger.warning("PyTables is not installed: writing %s as parquet instead of HDF5", group)
            filename = filename.with_suffix(".parquet")
    if filename.suffix not in (".parquet", ".csv"):
        filename = filename.with_suffix(".parquet")
    target = filename if group == "dynamics" else filename.with_suffix(f".{group}{filename.suffix}")
    return ("csv" if target.suffix == ".csv" else "parquet"), target


def _write_table(df: pandas.DataFrame, filename: Path, group: str, append: bool) -> None:
    """Write (``append=False``: start) or extend a table.  Appending never re-reads what is already on disk:
    HDF5 tables and CSV files are extended in place, parquet files one row group per call."""
    kind, target = _table_target(filename, group)
    if kind == "hdf":
        df.to_hdf(target, key=group, format="table", append=append)
        return
    if kind == "csv":
        df.to_csv(target, mode="a" if append else "w", header=not append, index=False)
        return
    import pyarrow as pa
    import pyarrow.parquet as pq

    key = str(target)
    table = pa.Table.from_pandas(df, preserve_index=False)
    writer = _parquet_writers.get(key)
    if writer is not None and not append:
        writer.close()
        writer = None
    if writer is None:
        if append and target.exists():
            # the file was closed by an earlier run: its rows move into the new writer once, then row groups follow
            old = pq.read_table(target)
            writer = pq.ParquetWriter(target, old.schema)
            writer.write_table(old)
        else:
            writer = pq.ParquetWriter(target, table.schema)
        _parquet_writers[key] = writer
    if table.schema != writer.schema:
        table = table.cast(writer.schema)
    writer.write_table(table)


def close_tables(filename: Optional[Path] = None) -> None:
    """Finish the parquet files written through :func:`_write_table` (all of them, or those of ``filename``)."""
    prefix = None if filename is None else str(Path(filename).with_suffix(""))
    for key in list(_parquet_writers):
        if prefix is None or key.startswith(prefix):
            _parquet_writers.pop(key).close()


def write_molecular_relaxations(engine, outfile: Path, temperature: float, pressure: float, keys=None) -> int:
    """The ``molecular_relaxations`` table (reference ``_read.py:177-190``: one row per (keyframe, molecule) with
    the passage time of every tracker), streamed one origin at a time: origins x N rows never exist at once (1 M
    molecules x 200 origins would be 2e8 rows).  Returns the number of rows written."""
    keys = list(engine.origins.keys()) if keys is None else list(keys)
    rows = 0
    molecule = np.arange(engine.n, dtype=np.int64)
    for i, key in enumerate(keys):
        columns = {"keyframe": np.full(engine.n, key), "molecule": molecule}
        columns.update(engine.summary(key))
        columns["temperature"] = np.full(engine.n, temperature, dtype=np.float64)
        columns["pressure"] = np.full(engine.n, pressure, dtype=np.float64)
        _write_table(pandas.DataFrame(columns), Path(outfile), "molecular_relaxations", append=i > 0)
        rows += engine.n
    return rows


class WriteCache:
    """Rows buffered in memory and flushed to ``filename`` every 8192 (reference ``_read.py:31-79``)."""

    def __init__(self, filename: Optional[Path] = None, group: str = "dynamics", cache_multiplier: int = 1, to_append: bool = False):
        if group is None:
            raise ValueError("Group can not be None.")
        self._filename = filename
        self.group = group
        self.cache_multiplier = cache_multiplier
        self.to_append = to_append
        self._cache: List[Any] = []
        self._blocks: List[Dict[str, Any]] = []  # column blocks (one per frame) from append_table
        self._block_rows = 0
        self._cache_default = 8192
        self._emptied_count = 0

    @property
    def _cache_size(self) -> int:
        return self.cache_multiplier * self._cache_default

    def append(self, item: Any) -> None:
        if self._cache and len(self._cache) == self._cache_size:
            self.flush()
            self._emptied_count += 1
        self._cache.append(item)

    def append_table(self, table: Dict[str, Any]) -> None:
        """All rows of one frame at once, as columns (what ``StepResult.table`` returns): the per-row dicts of
        ``append`` cost more than the kernels of a 32 k-molecule frame."""
        count = len(table["keyframe"])
        if self._filename is not None and self._block_rows + count > self._cache_size and (self._blocks or self._cache):
            self.flush()
            self._emptied_count += 1
        self._blocks.append(table)  # (a DataFrame per frame would cost more than the frame's kernels)
        self._block_rows += count

    def flush(self) -> None:
        assert self.filename is not None
        _write_table(self.to_dataframe(), self.filename, self.group, self.to_append)
        self.to_append = True
        self._cache.clear()
        self._blocks.clear()
        self._block_rows = 0

    def close(self) -> None:
        """Finish the output file (a parquet file becomes readable when its writer is closed; HDF5 / CSV need nothing)."""
        if self._filename is not None:
            close_tables(self.filename)

    @property
    def filename(self) -> Optional[Path]:
        return None if self._filename is None else Path(self._filename)

    def __len__(self) -> int:
        return self._cache_size * self._emptied_count + len(self._cache) + self._block_rows

    def to_dataframe(self):
        if not self._blocks:
            return pandas.DataFrame.from_records(self._cache)
        columns = list(self._blocks[0].keys())
        merged = pandas.DataFrame({
            name: np.concatenate([np.asarray(b[name]) for b in self._blocks]) for name in columns
        })
        if self._cache:
            merged = pandas.concat([pandas.DataFrame.from_records(self._cache), merged], ignore_index=True)
        return merged


def process_frames(
    file_iterator: FileIterator,
    wave_number: float,
    *,
    mol_relaxations: List[Dict[str, Any]] = None,
    scattering_function: bool = False,
    temperature: float = np.nan,
    pressure: float = np.nan,
    sink=None,
    table_sink=None,
    device=None,
    struct_mode: str = "reference",
    pipeline_depth: int = 3,
):
    """Run the hot loop of ``process_file`` (reference ``_read.py:128-168``) over any iterator of
    ``(indexes, frame)`` pairs.  Rows are passed to ``sink(row)`` one dict at a time like the reference builds
    them, or -- ``table_sink`` -- one column block per frame (``{quantity: (A,) array}``, the fast form);
    returns the engine (per-origin passage times stay on the device until ``engine.summary(key)`` is called)."""
    engine = None
    box_warned = False
    pending = deque()

    # table_sink: results are finalised in batches -- the closed forms of compute_all are ~15 numpy expressions,
    # which cost as much per call for one frame as for a hundred
    batch = []  # (keys, timediffs, sums, isf or None)
    batch_rows = 0

    def flush_batch():
        nonlocal batch_rows
        if not batch:
            return
        items, batch_rows = list(batch), 0
        batch.clear()

        def emit(part):
            sums = np.concatenate([it[2] for it in part])
            timediffs = np.concatenate([np.asarray(it[1], dtype=np.int64) for it in part])
            table = engine.finalize_table(sums, timediffs)  # ValueError: a row with non-unit quaternions
            if part[0][3] is not None:
                table["scattering_function"] = np.concatenate([it[3] for it in part]) / engine.n
            keys = [k for it in part for k in it[0]]
            table["keyframe"] = np.asarray(keys) if all(isinstance(k, (int, np.integer)) for k in keys) else keys
            table["temperature"] = np.full(len(keys), temperature)
            table["pressure"] = np.full(len(keys), pressure)
            table_sink(table)

        try:
            emit(items)
        except (ValueError, RuntimeError):
            for it in items:  # find the offending frame(s); the others keep their rows (_read.py:154-156)
                try:
                    emit([it])
                except (ValueError, RuntimeError) as e:
                    logger.warning(e)

    def drain(limit):
        nonlocal batch_rows
        while len(pending) > limit:
            result = pending.popleft()
            if table_sink is not None:
                if not result.keys:
                    continue
                try:
                    sums = result.sums()
                except (ValueError, RuntimeError) as e:  # _read.py:154-156
                    logger.warning(e)
                    continue
                batch.append((result.keys, result.timediffs, sums, result._isf))
                batch_rows += len(result.keys)
                if len(batch) >= 256 or batch_rows >= 8192:
                    flush_batch()
                continue
            try:
                rows = result.rows()
            except (ValueError, RuntimeError) as e:  # _read.py:154-156
                logger.warning(e)
                continue
            for key, row in zip(result.keys, rows):
                row["keyframe"] = key
                row["temperature"] = temperature
                row["pressure"] = pressure
                sink(row)

    for indexes, frame in file_iterator:
        if frame.position.shape[0] == 0:
            logger.warning("Found malformed frame ... continuing")
            continue
        logger.info("Indexes for step %s: %s", frame.timestep, indexes)
        if engine is None:
            engine = DynamicsEngine(
                frame.position.shape[0],
                frame.box,
                wave_number,
                trackers=mol_relaxations if mol_relaxations is not None else "default",
                mol_sites=Trimer().positions,
                device=device,
                struct_mode=struct_mode,
                ring_depth=pipeline_depth + 1,
                compute_isf=scattering_function,
            )
        elif not box_warned and not np.array_equal(np.asarray(frame.box, dtype=np.float32)[:2], engine.box[:2]):
            # The reference builds every keyframe's Dynamics / Relaxations with the box of that keyframe's
            # own first frame (_read.py:135-141); this engine wraps every origin with the box of the first
            # frame of the file.  On constant-volume trajectories the two are the same; on NPT files the
            # minimum-image step of later origins uses a slightly different L (results differ only for
            # steps within |dL| of L/2).  Documented deviation.
            logger.warning(
                "box changes along the trajectory (%s -> %s): all origins are wrapped with the first frame's box",
                engine.box[:2], np.asarray(frame.box)[:2],
            )
            box_warned = True
        try:
            pending.append(engine.step(frame.timestep, frame.position, frame.orientation, list(indexes)))
        except (ValueError, RuntimeError) as e:
            logger.warning(e)
            continue
        drain(pipeline_depth - 1)
    drain(0)
    flush_batch()
    return engine


def process_file(
    infile: Path,
    wave_number: float,
    steps_max: Optional[int] = None,
    linear_steps: Optional[int] = None,
    keyframe_interval: int = 1_000_000,
    keyframe_max: int = 500,
    mol_relaxations: List[Dict[str, Any]] = None,
    outfile: Optional[Path] = None,
    scattering_function: bool = False,
) -> Optional[pandas.DataFrame]:
    """Read a trajectory and compute the dynamic quantities of every time-origin.

    Same arguments and results as reference ``_read.py:82-195``: returns the ``dynamics`` DataFrame
    when ``outfile`` is None, otherwise writes the ``dynamics`` and ``molecular_relaxations`` tables
    and returns None.
    """
    assert infile is not None
    infile = Path(infile)
    logger.info("Processing %s", infile)
    start_time = datetime.datetime.now()

    sim_variables = get_filename_vars(infile)
    temperature = np.nan if sim_variables.temperature is None else float(sim_variables.temperature)
    pressure = np.nan if sim_variables.pressure is None else float(sim_variables.pressure)
    dataframes = WriteCache(filename=outfile) if outfile is not None else WriteCache()

    if infile.suffix == ".gsd":
        file_iterator = read_gsd_trajectory(
            infile=infile,
            steps_max=steps_max,
            linear_steps=linear_steps,
            keyframe_interval=keyframe_interval,
            keyframe_max=keyframe_max,
        )
    elif infile.suffix == ".lammpstrj":  # reference _read.py:121-122
        file_iterator = read_lammps_trajectory(infile, steps_max)
    else:
        raise ValueError(f"unsupported trajectory format {infile.suffix!r}")

    engine = process_frames(
        file_iterator,
        wave_number,
        mol_relaxations=mol_relaxations,
        scattering_function=scattering_function,
        temperature=temperature,
        pressure=pressure,
        table_sink=dataframes.append_table,
    )
    logger.info("Finished processing %s, took %s", infile, datetime.datetime.now() - start_time)

    if outfile is not None:
        dataframes.flush()
        if engine is not None and engine.trackers:
            write_molecular_relaxations(engine, Path(outfile), temperature, pressure)
        close_tables(Path(outfile))
        return None

    df = dataframes.to_dataframe()
    assert len(df) > 0
    return df


def open_trajectory(filename: Path, progressbar=None, frame_interval=1) -> Iterator[Frame]:
    """Frames of a GSD (skipping corrupt ones) or ``.lammpstrj`` trajectory (reference ``_read.py:198-241``)."""
    filename = Path(filename)
    if filename.suffix == ".lammpstrj":
        yield from parse_lammpstrj(filename)
        return
    if filename.suffix != ".gsd":
        raise ValueError(f"unsupported trajectory format {filename.suffix!r}")
    with open_gsd(filename) as trj:
        for index in range(0, len(trj), frame_interval):
            try:
                yield HoomdFrame(trj[index])
            except RuntimeError:
                logger.info("Found corrupt frame at index %s continuing", index)
                continue
