"""Argument plumbing of the ``comp-dynamics`` command (reference ``test/argparse_test.py:33-124``): the
compute entry point is replaced by a recorder, like the reference's ``dummy_process_file``."""

import pytest
from click.testing import CliRunner


@pytest.fixture()
def cli(monkeypatch):
    from sdanalysis_b200 import main

    calls = []
    monkeypatch.setattr(main, "process_file", lambda **kwargs: calls.append(kwargs))
    return main, calls


def test_version(cli):
    main, _ = cli
    from sdanalysis_b200 import __version__

    result = CliRunner().invoke(main.sdanalysis, ["--version"])
    assert result.exit_code == 0 and __version__ in result.output


@pytest.mark.parametrize("arg", ["-v", "-vv", "-vvv", "--verbose"])
def test_verbose(cli, arg, tmp_path):
    main, calls = cli
    infile = tmp_path / "traj.gsd"
    infile.write_bytes(b"")
    result = CliRunner().invoke(main.sdanalysis, [arg, "comp-dynamics", str(infile), str(tmp_path / "out.h5")])
    assert result.exit_code == 0, result.output
    assert len(calls) == 1


def test_missing_infile_is_a_usage_error(cli, tmp_path):
    main, calls = cli
    result = CliRunner().invoke(main.sdanalysis, ["comp-dynamics", "non-existent.file", str(tmp_path / "out.h5")])
    assert result.exit_code == 2 and not calls


def test_defaults_reach_process_file(cli, tmp_path):
    main, calls = cli
    infile = tmp_path / "traj.gsd"
    infile.write_bytes(b"")
    outfile = tmp_path / "sub" / "out.h5"
    result = CliRunner().invoke(main.sdanalysis, ["comp-dynamics", str(infile), str(outfile)])
    assert result.exit_code == 0, result.output
    (kwargs,) = calls
    # reference main.py:63-103 defaults
    assert kwargs["infile"] == infile and kwargs["outfile"] == outfile and outfile.parent.is_dir()
    assert kwargs["wave_number"] == 0.3 and kwargs["linear_steps"] == 100
    assert kwargs["keyframe_interval"] == 1_000_000 and kwargs["keyframe_max"] == 500
    assert "steps_max" not in kwargs and kwargs["mol_relaxations"] is None and kwargs["scattering_function"] is False


def test_options_and_aliases(cli, tmp_path):
    main, calls = cli
    infile = tmp_path / "traj.gsd"
    infile.write_bytes(b"")
    relax = tmp_path / "relax.yaml"
    relax.write_text("- name: tau_X\n  threshold: 0.5\n  last_passage: true\n")
    result = CliRunner().invoke(
        main.sdanalysis,
        ["-s", "1000", "--gen-steps", "10", "--max-gen", "7", "--wave-number", "2.9", "--molecule", "trimer",
         "comp-dynamics", "--linear-dynamics", "--scattering-function", "--mol-relaxations", str(relax),
         str(infile), str(tmp_path / "out.h5")],
    )
    assert result.exit_code == 0, result.output
    (kwargs,) = calls
    assert kwargs["steps_max"] == 1000 and kwargs["keyframe_interval"] == 10 and kwargs["keyframe_max"] == 7
    assert kwargs["wave_number"] == 2.9 and kwargs["linear_steps"] is None and kwargs["scattering_function"] is True
    assert kwargs["mol_relaxations"] == [{"name": "tau_X", "threshold": 0.5, "last_passage": True}]
    assert "molecule" not in kwargs and "output" not in kwargs
