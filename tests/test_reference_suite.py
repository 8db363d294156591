"""The reference's own pytest suite (smly/RiichiEnv tests/, unmodified) against riichienv_b200 — SURVEY §8 f1.

tests/refsuite/expected.txt lists the outcome of every reference test on the oracle backend and on the kernel code, with a
reason for everything that is not a pass.  Here the suite is run again and must reproduce the list:
  CPU (`-m "not gpu"`): the oracle backend and the host compile of the device sources;
  GPU (`-m gpu`):       the product through the C ABI — must reproduce the kernel column.
The suite is read from /root/reference/tests or from the git-ignored copy __graft_entry__.build() leaves in baseline/_ref/tests."""
import os

import pytest

from tests.refsuite import run as R

HERE = os.path.dirname(os.path.abspath(__file__))


def _expected():
    out = {}
    for line in open(os.path.join(HERE, "refsuite", "expected.txt")):
        if line.startswith("#") or not line.strip():
            continue
        p = [x.strip() for x in line.split("|")]
        out[p[0]] = (p[1], p[2], p[3] if len(p) > 3 else "")
    return out


def _check(backend, column):
    suite = R.find_suite()
    if suite is None:
        pytest.skip("reference test suite not present (neither /root/reference/tests nor baseline/_ref/tests)")
    got, log = R.run(backend, suite)
    exp = _expected()
    wrong = [f"{tid}: expected {exp[tid][column]}, got {got.get(tid, '-')}" for tid in exp if got.get(tid, "-") != exp[tid][column]]
    wrong += [f"{tid}: not in expected.txt (got {got[tid]})" for tid in got if tid not in exp]
    assert not wrong, "\n".join(wrong[:40]) + "\n" + log[-3000:]
    # every entry that is not a pass carries a reason
    assert all(v[2] for v in exp.values() if (v[0], v[1]) != ("pass", "pass"))
    assert sum(1 for v in exp.values() if v[column] == "pass") >= 210


def test_reference_suite_oracle():
    _check("oracle", 0)


def test_reference_suite_kernel_hostsim():
    _check("hostsim", 1)


@pytest.mark.gpu
def test_reference_suite_gpu():
    _check("gpu", 1)
