"""CPU tests of the command line's host logic (reference: src/main.cpp:484-672): argument validation, the messages the
reference prints, dir->dir naming with the collision rename, -j parsing.  No GPU work happens before validation ends."""
import os

import pytest

import realsr_ncnn_vulkan_b200.cli as cli


def run(argv, capsys):
    rc = cli.main(argv)
    return rc, capsys.readouterr().err


def test_validation_messages(tmp_path, capsys):
    f = str(tmp_path / "a.png")
    open(f, "wb").close()
    assert run([], capsys)[0] == -1
    assert "invalid scale argument" in run(["-i", f, "-o", f, "-s", "2"], capsys)[1]
    assert "invalid tilesize argument" in run(["-i", f, "-o", f, "-t", "16"], capsys)[1]
    assert "invalid tilesize argument" in run(["-i", f, "-o", f, "-t", "200,200", "-g", "0"], capsys)[1]
    assert "invalid thread count argument" in run(["-i", f, "-o", f, "-j", "0:2:2"], capsys)[1]
    assert "invalid jobs_proc thread count argument" in run(["-i", f, "-o", f, "-j", "1:2,2,2:2", "-g", "0,1"], capsys)[1]
    assert "invalid outputpath extension type" in run(["-i", f, "-o", str(tmp_path / "o.bmp")], capsys)[1]
    assert "unknown model dir type" in run(["-i", f, "-o", f, "-m", "models-foo"], capsys)[1]
    assert "must be either file or directory" in run(["-i", str(tmp_path), "-o", f], capsys)[1]
    assert "no CPU mode" in run(["-i", f, "-o", f, "-g", "-1"], capsys)[1]


def test_jobs_parsing():
    assert cli.parse_jobs("4:8:4") == (4, [8], 4)
    assert cli.parse_jobs("1:2,2,2:2") == (1, [2, 2, 2], 2)
    with pytest.raises(ValueError):
        cli.parse_jobs("4:8")


def test_directory_plan_and_collision_rename(tmp_path, capsys):
    ind, outd = tmp_path / "in", tmp_path / "out"
    ind.mkdir(); outd.mkdir()
    for n in ("b.png", "a.jpg", "a.png", "c.webp"):
        (ind / n).write_bytes(b"")
    ins, outs = cli.plan_files(str(ind), str(outd), "png")
    assert [os.path.basename(p) for p in ins] == ["a.jpg", "a.png", "b.png", "c.webp"]
    assert [os.path.basename(p) for p in outs] == ["a.png", "a.png.png", "b.png", "c.png"]
    assert "both a.png and a.jpg output a.png ! a.png will output a.png.png" in capsys.readouterr().err
