"""Runs the reference's own pytest suite (tests/ of smly/RiichiEnv, unmodified) against riichienv_b200.

The suite is not part of this repository: it is read from $RV_REF_TESTS, /root/reference/tests (the authoring
container) or baseline/_ref/tests (a git-ignored copy that __graft_entry__.build() makes where the reference exists, so
that it travels to the GPU box with the snapshot).  `riichienv` resolves to tests/refsuite/riichienv, a re-export of
riichienv_b200; RV_REFSUITE_BACKEND picks what executes the game logic (gpu = the product; oracle / hostsim = the CPU
checkers).  Returns {test id: outcome} from the junit report.
"""
import os
import subprocess
import sys
import tempfile
import xml.etree.ElementTree as ET

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def find_suite():
    for d in (os.environ.get("RV_REF_TESTS"), "/root/reference/tests", os.path.join(ROOT, "baseline", "_ref", "tests")):
        if d and os.path.isdir(os.path.join(d, "env")):
            return d
    return None


def run(backend, suite=None, timeout=1800):
    suite = suite or find_suite()
    if suite is None:
        raise FileNotFoundError("reference test suite not found (RV_REF_TESTS, /root/reference/tests, baseline/_ref/tests)")
    env = dict(os.environ)
    env["RV_REFSUITE_BACKEND"] = backend
    env["PYTHONPATH"] = os.pathsep.join([HERE, ROOT] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    with tempfile.TemporaryDirectory() as tmp:
        xml = os.path.join(tmp, "report.xml")
        # rootdir = the suite's parent so that ids read tests/env/...; no cache / bytecode is written into the suite
        cmd = [sys.executable, "-B", "-m", "pytest", suite, "-q", "-p", "no:cacheprovider", "--continue-on-collection-errors",
               "--timeout", "300", "--junitxml", xml, "-o", "junit_family=xunit1", "--rootdir", os.path.dirname(suite),
               "-c", os.devnull]
        p = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True, timeout=timeout)
        if not os.path.exists(xml):
            raise RuntimeError(f"pytest produced no report:\n{p.stdout[-4000:]}\n{p.stderr[-4000:]}")
        out = {}
        for case in ET.parse(xml).getroot().iter("testcase"):
            cls, name = case.get("classname", ""), case.get("name", "")
            tid = f"{cls}::{name}" if name else cls
            if case.find("failure") is not None:
                res = "fail"
            elif case.find("error") is not None:
                res = "error"
            elif case.find("skipped") is not None:
                res = "skip"
            else:
                res = "pass"
            out[tid] = res
        return out, p.stdout


if __name__ == "__main__":
    backend = sys.argv[1] if len(sys.argv) > 1 else "oracle"
    res, log = run(backend)
    for k in sorted(res):
        print(f"{res[k]:5s} {k}")
    n = {o: sum(1 for v in res.values() if v == o) for o in ("pass", "fail", "error", "skip")}
    print(f"# backend={backend} {n}", file=sys.stderr)
