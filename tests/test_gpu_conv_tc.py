"""tcgen05 / TMA conv engine vs the generic CUDA-core engine (and ATen) on every tensor-core geometry of the path."""
import pytest

pytestmark = pytest.mark.gpu

from tests import tc_cases  # noqa: E402


@pytest.mark.parametrize("idx", range(len(tc_cases.CASES)), ids=lambda i: "case%d" % i)
def test_tc_engine_matches_generic(idx):
    res = tc_cases.run_case(idx, vs_cpu=(idx % 3 == 0))
    bad = tc_cases.check(res)
    assert not bad, "%s: out of tolerance %s in %s" % (res["case"], bad, res)


def test_tc_pair_mode_matches_generic():
    """CTA-pair kernels (tcgen05 cta_group::2) on every geometry they take.  Runs in a child process with
    NEMAR_TC_PAIR=1 (a device trap must not poison this process's context)."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, json; sys.path.insert(0, %r); from tests import tc_cases as t\n"
            "for i in t.PAIR_CASES:\n"
            "    r = t.run_case(i, True); r['bad'] = t.check(r); print('RES ' + json.dumps(r), flush=True)\n") % root
    env = dict(os.environ, NEMAR_TC_PAIR="1")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    res = [json.loads(l[4:]) for l in p.stdout.splitlines() if l.startswith("RES ")]
    assert p.returncode == 0 and len(res) == len(tc_cases.PAIR_CASES), (p.returncode, p.stderr[-1500:])
    bad = [r for r in res if r["bad"]]
    assert not bad, bad


def test_tc_resident_patch_mode_matches_generic():
    """Resident-patch kernel (tap-shifted UMMA windows over one TMA patch; opt-in NEMAR_TC_RP3=1) on every stride-1
    k x k geometry it takes.  Child process (a device trap must not poison this one).  First green on a B200 in
    round 2 (profiles/r02_rp3_cases.txt)."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, json; sys.path.insert(0, %r); from tests import tc_cases as t\n"
            "for i in t.RP3_CASES:\n"
            "    r = t.run_case(i, True); r['bad'] = t.check(r); print('RES ' + json.dumps(r), flush=True)\n") % root
    env = dict(os.environ, NEMAR_TC_RP3="1")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    res = [json.loads(l[4:]) for l in p.stdout.splitlines() if l.startswith("RES ")]
    assert p.returncode == 0 and len(res) == len(tc_cases.RP3_CASES), (p.returncode, p.stderr[-1500:])
    bad = [r for r in res if r["bad"]]
    assert not bad, bad
