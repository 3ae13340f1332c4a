"""pip64, the command-line front end (piplib_b200/csrc/pip_cli.cpp), against byte-exact transcripts of
the reference's own tool (tests/golden/cli_text.json, made by tools/make_cli_golden.py from
source/maind.c): every test/*.dat, all of them in one file (one device batch), -d, a syntax error,
a fatal verdict's exit code, and tab_get's skip-to-']' quirk after an empty context."""
import os
import subprocess

import pytest

from conftest import load_golden
from oracle import pyoracle as po

CASES = load_golden("cli_text.json")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIP64 = os.path.join(ROOT, "piplib_b200", "bin", "pip64")


@pytest.fixture(scope="module")
def pip64():
    from piplib_b200 import build
    build.build()
    assert os.path.exists(PIP64)
    return PIP64


def test_lexer_matches_the_dat_grammar(pip64):
    """no GPU needed: --lex-only dumps what the front end read"""
    seen = 0
    for c in CASES:
        if c["args"] != ["-s"] or c["name"] in ("all-in-one-file",) or "-" in c["name"]:
            continue
        want = po.parse_dat(c["input"])
        r = subprocess.run([pip64, "--lex-only"], input=c["input"].encode("latin-1"), capture_output=True)
        assert r.returncode == 0
        lines = r.stdout.decode("latin-1").split("\n")
        head = lines[0].split()
        assert head[1] == "error=0"
        assert [int(x) for x in head[3:9]] == [want[k] for k in ("nvar", "nparm", "ni", "nc", "bigparm", "nq")]
        assert [int(x) for x in lines[1].split()] == [v for row in want["tab"] for v in row]
        assert [int(x) for x in lines[2].split()] == [v for row in want["ctx"] for v in row]
        assert int(head[2].split("=")[1]) == len(want["comment"])
        seen += 1
    assert seen >= 30


def test_lexer_reports_syntax_errors_and_keeps_going(pip64):
    c = [x for x in CASES if x["name"] == "syntax-error-then-problem"][0]
    r = subprocess.run([pip64, "--lex-only"], input=c["input"].encode("latin-1"), capture_output=True)
    heads = [ln for ln in r.stdout.decode().split("\n") if ln.startswith("problem")]
    assert len(heads) == 2 and "error=1" in heads[0] and "error=0" in heads[1]


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_transcripts(pip64, c):
    r = subprocess.run([pip64] + c["args"], input=c["input"].encode("latin-1"), capture_output=True, timeout=300)
    assert r.stdout.decode("latin-1") == c["stdout"]
    assert r.returncode == c["rc"]
