"""tcgen05 / TMA conv engine vs the generic CUDA-core engine (and ATen) on every tensor-core geometry of the path."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

from tests import tc_cases  # noqa: E402


@pytest.mark.parametrize("idx", range(len(tc_cases.CASES)), ids=lambda i: "case%d" % i)
def test_tc_engine_matches_generic(idx):
    """Default kernel selection (256-wide tiles for 256-channel destinations, resident-patch kernel for stride-1 k x k
    layers with <= 64 output channels, CTA-pair weight gradients, tiled gather elsewhere)."""
    res = tc_cases.run_case(idx, vs_cpu=(idx % 3 == 0))
    bad = tc_cases.check(res)
    assert not bad, "%s: out of tolerance %s in %s" % (res["case"], bad, res)


def _run_variant(env, cases):
    """every case in ONE child process with the engine switches of `env` (read once per process; a device trap must
    not poison this process's context)"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, json; sys.path.insert(0, %r); from tests import tc_cases as t\n"
            "for i in %r:\n"
            "    r = t.run_case(i, True); r['bad'] = t.check(r); print('RES ' + json.dumps(r), flush=True)\n") % (root, list(cases))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=dict(os.environ, **env))
    res = [json.loads(l[4:]) for l in p.stdout.splitlines() if l.startswith("RES ")]
    assert p.returncode == 0 and len(res) == len(cases), (p.returncode, p.stderr[-1500:])
    bad = [r for r in res if r["bad"]]
    assert not bad, bad


def test_tc_pair_mode_matches_generic():
    """CTA-pair kernels (tcgen05 cta_group::2) in the gather kernel too (NEMAR_TC_PAIR=1) on every geometry they take."""
    _run_variant({"NEMAR_TC_PAIR": "1"}, tc_cases.PAIR_CASES)


def test_tc_tiled_kernels_match_generic():
    """The kernels the defaults replace: 128-wide destination tiles (NEMAR_TC_WIDE=0) and the tiled gather on the
    small-channel stride-1 layers (NEMAR_TC_RP3=0) — they remain the fallback for geometries the newer kernels decline."""
    _run_variant({"NEMAR_TC_WIDE": "0", "NEMAR_TC_RP3": "0"}, range(len(tc_cases.CASES)))
