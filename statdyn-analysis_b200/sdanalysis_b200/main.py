"""``sdanalysis comp-dynamics`` on the B200 path: the command-line surface of reference ``main.py:44-165``
for the one sub-command that drives the hot loop.

    python -m sdanalysis_b200 [group options] comp-dynamics [--linear-dynamics] [--scattering-function]
                                                            [--mol-relaxations FILE] INFILE OUTFILE

Option names, defaults and the way they reach ``read.process_file`` are the reference's
(``main.py:63-111,140-165``).  Two defects of the reference are not reproduced: ``--molecule`` and
``--output`` are accepted but not forwarded (the reference passes them on to ``process_file``, which has
no such arguments), and ``--mol-relaxations`` is read as a YAML document (the reference hands the file
*name* to ``yaml.parse``; SURVEY trap T9).  ``comp-relaxations`` and ``figure`` are post-processing /
plotting and are not part of this package.
"""

import logging
from pathlib import Path

import click

from . import __version__
from .read import process_file

logger = logging.getLogger(__name__)

MOLECULE_OPTIONS = ("trimer", "disc", "sphere", "dimer")
_PROCESS_FILE_OPTIONS = ("wave_number", "steps_max", "linear_steps", "keyframe_interval", "keyframe_max")


def _verbosity(_, __, value) -> None:
    levels = {0: "WARNING", 1: "INFO", 2: "DEBUG"}
    logging.basicConfig(level=levels.get(value, "DEBUG"))


def _file_logger(_, __, value) -> None:
    if value:
        logging.basicConfig(filename=value, level=logger.level)


@click.group()
@click.version_option(__version__)
@click.option("-v", "--verbose", count=True, default=0, callback=_verbosity, expose_value=False, is_eager=True,
              help="Increase debug level")
@click.option("--log-file", callback=_file_logger, expose_value=False, is_eager=True,
              help="Set output file for logging information.")
@click.option("-s", "--num-steps", "--steps-max", "steps_max", type=int, help="Maximum number of steps for analysis")
@click.option("--keyframe-interval", "--gen-steps", "keyframe_interval", type=int, default=1_000_000,
              help="Steps between key frames in simulation")
@click.option("--linear-steps", type=int, default=100, help="Number of steps between exponential increase.")
@click.option("--keyframe-max", "--max-gen", "keyframe_max", type=int, default=500, help="Maximum number of key frames")
@click.option("--molecule", type=click.Choice(MOLECULE_OPTIONS), help="Molecule to use for simulation")
@click.option("--wave-number", type=float, default=0.3,
              help="This is the wave number corresponding to the maximum of the structure factor")
@click.option("-o", "--output", type=click.Path(file_okay=False, dir_okay=True),
              help="Location to save all output files; required to be a directory.")
@click.pass_context
def sdanalysis(ctx, **kwargs) -> None:
    """Dynamics analysis of a trajectory on B200 GPUs."""
    ctx.obj = {key: val for key, val in kwargs.items() if val is not None}


@sdanalysis.command("comp-dynamics")
@click.pass_obj
@click.option("--mol-relaxations", type=click.Path(exists=True, dir_okay=False),
              help="Path to a file defining all the molecular relaxations to compute.")
@click.option("--linear-dynamics", type=bool, default=False, is_flag=True,
              help="Flag to specify the configurations in a trajectory have linear steps between them.")
@click.option("--scattering-function", type=bool, default=False, is_flag=True,
              help="Calculate the values for the intermediate scattering function.")
@click.argument("infile", type=click.Path(exists=True, file_okay=True, dir_okay=False))
@click.argument("outfile", type=click.Path(file_okay=True, dir_okay=False))
def comp_dynamics(obj, mol_relaxations, linear_dynamics, scattering_function, infile, outfile) -> None:
    """Compute dynamic properties for a single input file."""
    outfile, infile = Path(outfile), Path(infile)
    outfile.parent.mkdir(parents=True, exist_ok=True)
    options = {key: obj[key] for key in _PROCESS_FILE_OPTIONS if key in (obj or {})}
    if linear_dynamics:
        options["linear_steps"] = None
    relaxations = None
    if mol_relaxations is not None:
        import yaml

        with open(mol_relaxations) as src:
            relaxations = yaml.safe_load(src)
    logger.debug("Processing: %s", infile)
    process_file(infile=infile, mol_relaxations=relaxations, outfile=outfile, scattering_function=scattering_function,
                 **options)


if __name__ == "__main__":
    sdanalysis()
