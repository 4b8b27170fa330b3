"""Bounds evidence without compute-sanitizer (closed on the GPU pool): the kernels are rebuilt with -DMCB_CANARY -- guard words
between the arrays of every per-env shared-memory record, planted when an env is loaded and verified when it is stored -- and
run through every controller, tier and reset path (tools/canary_workload.py).  A self-test proves the guards fire."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    from mycobotgym_b200 import _lib

    assert os.path.exists(_lib.CANARY_PATH), "build the canary library first: python -c 'import __graft_entry__ as g; g.build()'"
    env = dict(os.environ, MCB_LIB=_lib.CANARY_PATH, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", "canary_workload.py")], env=env, capture_output=True, text=True, timeout=900)


def test_no_guard_word_is_overwritten():
    r = _run({})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "CANARY_OK" in r.stdout and "MCB_CANARY:" not in r.stdout, r.stdout[-3000:]
    print(r.stdout)


def test_guards_detect_a_deliberate_overrun():
    r = _run({"MCB_CANARY_SELFTEST": "1"})
    assert "CANARY_HIT" in r.stdout and "MCB_CANARY: guard 7" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
