"""tcgen05 / TMA conv engine vs the generic CUDA-core engine (and ATen) on every tensor-core geometry of the path."""
import pytest

pytestmark = pytest.mark.gpu

from tests import tc_cases  # noqa: E402


@pytest.mark.parametrize("idx", range(len(tc_cases.CASES)), ids=lambda i: "case%d" % i)
def test_tc_engine_matches_generic(idx):
    res = tc_cases.run_case(idx, vs_cpu=(idx % 3 == 0))
    bad = tc_cases.check(res)
    assert not bad, "%s: out of tolerance %s in %s" % (res["case"], bad, res)
