"""The yacrd-compatible driver (yacrd_b200/csrc/cli.cpp <- reference src/main.rs:36-137, src/cli.rs:39-74).
GPU part replays the reference's own integration test for detection (tests/run.rs:96-117): run the binary on
tests/reads.paf and compare the report with tests/truth.yacrd as an unordered set of lines."""
import os
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")


def read_sorted_lines(path):
    with open(path) as fh:
        return sorted(l for l in fh.read().split("\n") if l)

CLI = os.path.join(REPO, "yacrd_b200", "yacrd-b200")


def run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=300)


def test_cli_version_help_and_argument_errors():
    assert os.path.exists(CLI), "build() must produce the driver"
    r = run("--version")
    assert r.returncode == 0 and r.stdout.startswith("yacrd 1.0.0 Magby")  # cli.rs:35
    r = run("--help")
    assert r.returncode == 0 and "--not-coverage" in r.stdout and "--coverage" in r.stdout
    r = run("-i", "x.paf")
    assert r.returncode == 2 and "--output" in r.stderr
    r = run("-i", "x.paf", "-o", "y", "-c", "abc")
    assert r.returncode == 2
    r = run("-i", "x.paf", "-o", "y", "scrubb", "-i", "a.fq")  # cli.rs:94-103: both are required
    assert r.returncode == 2 and "--output" in r.stderr
    r = run("-i", "x.paf", "-o", "y", "filter", "-i", "a.fq", "-o", "b.fq", "--bogus")
    assert r.returncode == 2


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_overlaps.paf", "c1_overlaps.m4"])
def test_cli_detection_matches_the_reference_truth(tmp_path, name):
    out = tmp_path / "out.yacrd"
    r = run("-i", os.path.join(GOLDEN, name), "-o", str(out), "--timing")
    assert r.returncode == 0, r.stderr
    assert read_sorted_lines(str(out)) == read_sorted_lines(os.path.join(GOLDEN, "c1_truth.sorted.yacrd"))
    assert "230 reads" in r.stderr
    # report as input (main.rs:43-45): classification re-run on the device, same lines come back
    out2 = tmp_path / "again.yacrd"
    r = run("-i", str(out), "-o", str(out2))
    assert r.returncode == 0, r.stderr
    assert read_sorted_lines(str(out2)) == read_sorted_lines(str(out))


@pytest.mark.gpu
@pytest.mark.parametrize("c,n,golden", [(4, 0.4, "c1_oracle_c4_n0.4.sorted.yacrd"), (3, 0.4, "c1_oracle_c3_n0.4.sorted.yacrd"),
                                        (1, 0.8, "c1_oracle_c1_n0.8.sorted.yacrd")])
def test_cli_presets_match_the_oracle(tmp_path, c, n, golden):
    out = tmp_path / "out.yacrd"
    r = run("-i", os.path.join(GOLDEN, "c1_overlaps.paf"), "-o", str(out), "-c", str(c), "-n", str(n), "-t", "4")
    assert r.returncode == 0, r.stderr
    assert read_sorted_lines(str(out)) == read_sorted_lines(os.path.join(GOLDEN, golden))


@pytest.mark.gpu
def test_cli_ondisk_mode_streams_chunks_and_gives_the_same_report(tmp_path):
    """`-d <prefix> --ondisk-buffer-size <bytes>` (cli.rs:61-70): here chunks of that many bytes of intervals go through the
    device on the streamed path's lanes (chunks hold whole reads, a multiple of 1024 of them: the golden PAF's 230 reads
    travel as one chunk); the report is the reference's."""
    out = tmp_path / "out.yacrd"
    r = run("-i", os.path.join(GOLDEN, "c1_overlaps.paf"), "-o", str(out), "-d", str(tmp_path / "ondisk"), "--ondisk-buffer-size", "8192")
    assert r.returncode == 0, r.stderr
    assert read_sorted_lines(str(out)) == read_sorted_lines(os.path.join(GOLDEN, "c1_truth.sorted.yacrd"))
    assert not os.path.exists(str(tmp_path / "ondisk")), "nothing is written to disk"
    r = run("-i", os.path.join(GOLDEN, "c1_overlaps.paf"), "-o", str(out), "-d", "x", "--ondisk-buffer-size", "abc")
    assert r.returncode == 2


@pytest.mark.gpu
def test_cli_errors_like_the_reference(tmp_path):
    r = run("-i", str(tmp_path / "missing.paf"), "-o", str(tmp_path / "o.yacrd"))
    assert r.returncode == 1 and "Can't open file" in r.stderr                       # error.rs CantReadFile
    bad = tmp_path / "reads.fasta"
    bad.write_text(">a\nACGT\n")
    r = run("-i", str(bad), "-o", str(tmp_path / "o.yacrd"))
    assert r.returncode == 1 and "fasta" in r.stderr                                  # error.rs CantRunOperationOnFile
    unk = tmp_path / "reads.txt"
    unk.write_text("x\n")
    r = run("-i", str(unk), "-o", str(tmp_path / "o.yacrd"))
    assert r.returncode == 1 and "Format detection" in r.stderr                       # error.rs UnableToDetectFileFormat
