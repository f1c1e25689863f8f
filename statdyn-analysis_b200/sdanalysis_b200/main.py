"""``sdanalysis comp-dynamics`` on the B200 path: the command-line surface of reference ``main.py:44-165``
for the one sub-command that drives the hot loop.

    python -m sdanalysis_b200 [group options] comp-dynamics [--linear-dynamics] [--scattering-function]
                                                            [--mol-relaxations FILE] INFILE OUTFILE

Option names, aliases, defaults and the way they reach ``read.process_file`` are the reference's
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

LOG_LEVELS = ("WARNING", "INFO", "DEBUG")
MOLECULES = ("trimer", "disc", "sphere", "dimer")
# group options that process_file understands (everything else stays on the command line)
FORWARDED = ("wave_number", "steps_max", "linear_steps", "keyframe_interval", "keyframe_max")


def _set_verbosity(ctx, param, count):
    logging.basicConfig(level=LOG_LEVELS[min(count, len(LOG_LEVELS) - 1)])


def _set_log_file(ctx, param, path):
    if path:
        logging.basicConfig(filename=path, level=logger.level)


# (flags and destination, click keyword arguments) of the option group shared by all sub-commands
GROUP_OPTIONS = (
    (("-v", "--verbose"), dict(count=True, default=0, callback=_set_verbosity, expose_value=False, is_eager=True,
                               help="more logging: -v INFO, -vv DEBUG")),
    (("--log-file",), dict(callback=_set_log_file, expose_value=False, is_eager=True, help="write the log to this file")),
    (("-s", "--num-steps", "--steps-max", "steps_max"), dict(type=int, help="last timestep to analyse")),
    (("--keyframe-interval", "--gen-steps", "keyframe_interval"),
     dict(type=int, default=1_000_000, help="timesteps between the starts of successive time-origins")),
    (("--linear-steps",), dict(type=int, default=100, help="linearly spaced steps per decade of the step schedule")),
    (("--keyframe-max", "--max-gen", "keyframe_max"), dict(type=int, default=500, help="largest number of time-origins")),
    (("--molecule",), dict(type=click.Choice(MOLECULES), help="molecule of the simulation (accepted, not used)")),
    (("--wave-number",), dict(type=float, default=0.3, help="wave number of the first peak of the structure factor")),
    (("-o", "--output"), dict(type=click.Path(file_okay=False, dir_okay=True), help="output directory (accepted, not used)")),
)


def _options(table):
    def decorate(function):
        for flags, keywords in reversed(table):
            function = click.option(*flags, **keywords)(function)
        return function

    return decorate


@click.group()
@click.version_option(__version__)
@_options(GROUP_OPTIONS)
@click.pass_context
def sdanalysis(ctx, **values) -> None:
    """Dynamics analysis of a trajectory on B200 GPUs."""
    ctx.obj = {name: value for name, value in values.items() if value is not None}


@sdanalysis.command("comp-dynamics")
@click.option("--mol-relaxations", type=click.Path(exists=True, dir_okay=False),
              help="YAML file with the molecular relaxation definitions (name, threshold, ...)")
@click.option("--linear-dynamics", is_flag=True, default=False,
              help="the frames are linearly spaced: every frame advances every time-origin")
@click.option("--scattering-function", is_flag=True, default=False,
              help="also compute the intermediate scattering function")
@click.argument("infile", type=click.Path(exists=True, dir_okay=False))
@click.argument("outfile", type=click.Path(dir_okay=False))
@click.pass_obj
def comp_dynamics(group_values, infile, outfile, mol_relaxations, linear_dynamics, scattering_function) -> None:
    """Dynamic quantities and molecular relaxation times of INFILE, written to OUTFILE."""
    source, target = Path(infile), Path(outfile)
    target.parent.mkdir(parents=True, exist_ok=True)
    arguments = {name: group_values[name] for name in FORWARDED if name in (group_values or {})}
    if linear_dynamics:
        arguments["linear_steps"] = None  # what selects the dense mode in process_file
    definitions = None
    if mol_relaxations is not None:
        import yaml

        definitions = yaml.safe_load(Path(mol_relaxations).read_text())
    logger.debug("Processing: %s", source)
    process_file(infile=source, outfile=target, mol_relaxations=definitions, scattering_function=scattering_function,
                 **arguments)


if __name__ == "__main__":
    sdanalysis()
