"""CLI behaviour that needs no GPU: argument and input errors keep the reference's "message on stderr, exit 1"."""
import subprocess

import pytest

import cuda_satabsearch_b200 as S

CLI = S.CLI_PATH


def run(args, stdin=b""):
    return subprocess.run([str(CLI)] + args, input=stdin, capture_output=True, timeout=60)


@pytest.mark.skipif(not CLI.exists(), reason="CLI not built")
def test_cli_errors_exit_1():
    p = run(["-c"], b"x\nT T F\n")
    assert p.returncode == 1 and b"-c" in p.stderr and p.stdout == b""          # no CPU search path
    p = run([], b"/nonexistent/db.ascii\nT T F\nA          1\ne  \n 0.000 \n")
    assert p.returncode == 1 and b"ERROR opening db file" in p.stderr
    p = run([], b"")
    assert p.returncode == 1 and b"ERROR" in p.stderr
    p = run(["-r", "0"], b"x\nT T F\n")
    assert p.returncode == 1
    p = run(["-Z"])
    assert p.returncode == 1 and b"Usage" in p.stderr
    p = run(["-q", "/nonexistent/db.ascii"], b"d1ubia_\n")
    assert p.returncode == 1 and b"ERROR opening db file" in p.stderr
