#!/usr/bin/env python
"""tests/golden/cli_text.json: byte-exact transcripts of the reference's command-line tool
(oracle/_ref/pip_dp = source/maind.c + the five library files, built by `make -C oracle refcli`) on
  * every test/*.dat file (the `make check` of the reference is this command diffed with the .ll),
  * all of them concatenated into one multi-problem file (the batch mode of our pip64), which also
    exercises tab_get's skip to the next ']' after an empty context,
  * the -z (simplify) and -d (deepest cut) switches, a syntax error, a fatal verdict (exit code).
Run in the build container only; the tests read the JSON."""
import glob
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PIP = os.path.join(ROOT, "oracle", "_ref", "pip_dp")
SKIP = {"boulet", "bouleti"}          # seconds of CPU each, covered cell for cell elsewhere


def run(args, text):
    r = subprocess.run([PIP] + args, input=text.encode("latin-1"), capture_output=True, timeout=120)
    return r.returncode, r.stdout.decode("latin-1")


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "refcli"])
    cases, texts = [], []
    for f in sorted(glob.glob(REF + "/test/*.dat")):
        name = os.path.basename(f)[:-4]
        if name in SKIP:
            continue
        text = open(f, encoding="latin-1").read()
        texts.append(text)
        rc, out = run(["-s"], text)
        cases.append(dict(name=name, args=["-s"], input=text, rc=rc, stdout=out))
    rc, out = run(["-s"], "".join(texts))
    cases.append(dict(name="all-in-one-file", args=["-s"], input="".join(texts), rc=rc, stdout=out))
    for name in ("max", "rairoi", "lineri", "pairi"):
        text = open("%s/test/%s.dat" % (REF, name), encoding="latin-1").read()
        for args in (["-s", "-d"],):
            rc, out = run(args, text)
            cases.append(dict(name=name + " ".join(args), args=args, input=text, rc=rc, stdout=out))
    # (with nq = 1 the reference dereferences the NULL tableau before checking it: source/maind.c:189-191)
    bad = "((broken) 2 0 2 0 -1 0 (#[1 0 x] #[1 2 -3]) ())\n" + open(REF + "/test/test3i.dat", encoding="latin-1").read()
    rc, out = run(["-s"], bad)
    cases.append(dict(name="syntax-error-then-problem", args=["-s"], input=bad, rc=rc, stdout=out))
    chal = open(REF + "/test/challenges/pipFile_0", encoding="latin-1").read()
    rc, out = run(["-s"], open(REF + "/test/test2i.dat", encoding="latin-1").read() + chal)
    cases.append(dict(name="empty-context-swallows-next-problem", args=["-s"],
                      input=open(REF + "/test/test2i.dat", encoding="latin-1").read() + chal, rc=rc, stdout=out))
    rc, out = run(["-s"], chal)
    cases.append(dict(name="fatal-verdict-exit-code", args=["-s"], input=chal, rc=rc, stdout=out))
    mixed = open(REF + "/test/max.dat", encoding="latin-1").read() + chal
    rc, out = run(["-s"], mixed)
    cases.append(dict(name="fatal-after-good", args=["-s"], input=mixed, rc=rc, stdout=out))
    path = os.path.join(ROOT, "tests", "golden", "cli_text.json")
    json.dump(cases, open(path, "w"), separators=(",", ":"))
    print(len(cases), "cases,", os.path.getsize(path), "bytes; exit codes", sorted({c["rc"] for c in cases}))


if __name__ == "__main__":
    main()
