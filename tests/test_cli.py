"""python -m soundswallower_b200: the reference CLI's surface (ref: py/soundswallower/cli.py:139-171,
py/test/test_cli.py:30-99) plus --align-batch."""
import json
import os

import pytest

from conftest import DATA
from soundswallower_b200 import cli
from test_decoder_api import CLI_JSON, TEXT


def test_cli_parser_mirrors_the_reference():
    p = cli.make_argparse()
    a = p.parse_args(["--align", "x.txt", "--phone-align", "--model", "fr-fr", "-o", "out.json", "a.wav", "b.wav"])
    assert (a.align, a.phone_align, a.model, a.output, a.inputs) == ("x.txt", True, "fr-fr", "out.json", ["a.wav", "b.wav"])
    with pytest.raises(SystemExit):
        p.parse_args(["--align", "x.txt", "--fsg", "y.fsg"])       # mutually exclusive, as in the reference
    assert cli.model_path("en-us").endswith(os.path.join("model", "en-us")) and cli.model_path("/x/y") == "/x/y"
    assert cli.main([]) is None                                     # nothing to do


def test_cli_batch_list_parsing(tmp_path):
    (tmp_path / "t.txt").write_text("go forward ten meters\n")
    lst = tmp_path / "list.tsv"
    lst.write_text("# comment\n%s\tgo forward ten meters\nrel.wav\t@t.txt\n\n" % os.path.join(DATA, "goforward.raw"))
    files, texts = cli.read_batch_list(str(lst))
    assert files == [os.path.join(DATA, "goforward.raw"), str(tmp_path / "rel.wav")]
    assert texts == ["go forward ten meters"] * 2
    lst.write_text("no-tab-here\n")
    with pytest.raises(SystemExit):
        cli.read_batch_list(str(lst))


@pytest.mark.gpu
def test_cli_align_is_the_reference_cli_line(tmp_path, capsys):
    """soundswallower --align goforward.txt --phone-align goforward.raw."""
    cli.main(["--align", os.path.join(DATA, "goforward.txt"), "--phone-align", os.path.join(DATA, "goforward.raw")])
    assert capsys.readouterr().out == CLI_JSON
    out = tmp_path / "o.json"
    cli.main(["--align-text", TEXT["en-us"], "-o", str(out), os.path.join(DATA, "goforward.raw"),
              os.path.join(DATA, "goforward.raw")])
    lines = out.read_text().splitlines()
    assert len(lines) == 2 and lines[0] == lines[1]
    assert [w["t"] for w in json.loads(lines[0])["w"]] == ["<sil>", "go", "forward", "ten", "meters", "<sil>"]


@pytest.mark.gpu
def test_cli_align_batch_equals_one_file_at_a_time(tmp_path, capsys):
    wav = os.path.join(DATA, "goforward.raw")
    lst = tmp_path / "list.tsv"
    lst.write_text("%s\t%s\n%s\t@%s\n%s\tgo backward ten meters\n" % (wav, TEXT["en-us"], wav,
                                                                     os.path.join(DATA, "goforward.txt"), wav))
    cli.main(["--align-batch", str(lst), "--phone-align"])
    lines = capsys.readouterr().out.splitlines(keepends=True)
    assert len(lines) == 3
    assert lines[0] == CLI_JSON and lines[1] == CLI_JSON
    # a transcript that does not match still aligns *something* or yields null; never a crash
    assert lines[2] == "null\n" or json.loads(lines[2])["t"]


@pytest.mark.gpu
def test_cli_fsg_file_decodes(capsys):
    cli.main(["--fsg", os.path.join(DATA, "goforward.fsg"), os.path.join(DATA, "goforward.raw")])
    assert json.loads(capsys.readouterr().out)["t"] == "go forward ten meters"
