"""Golden fixtures (tests/golden/streams.json, written by tests/golden/make_golden.py): the oracle must keep
reproducing them (CPU), and the CUDA engine must reproduce them too (`-m gpu`) — a second, file-based pin
next to the live oracle-vs-engine comparison."""
import json
import os

import pytest

from tests.golden import make_golden

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "streams.json")))


@pytest.mark.parametrize("case", make_golden.CASES)
def test_oracle_reproduces_golden(case):
    assert make_golden.run_case(case) == GOLD[case]


@pytest.mark.gpu
@pytest.mark.parametrize("case", make_golden.CASES)
def test_engine_reproduces_golden(case):
    from rafting_b200 import engine
    assert make_golden.run_case(case, sut_factory=engine.Engine) == GOLD[case]
