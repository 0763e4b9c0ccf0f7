"""baseline/_ref must be the reference byte for byte (CPU test; only meaningful in the build container where
/root/reference exists)."""
import filecmp
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = "/root/reference", os.path.join(ROOT, "baseline", "_ref")


def test_shipped_reference_is_unmodified():
    if not os.path.isdir(SRC) or not os.path.isdir(DST):
        pytest.skip("needs /root/reference and baseline/_ref")
    n = 0
    for base in ("disvae", "utils"):
        for d, _, files in os.walk(os.path.join(SRC, base)):
            for f in files:
                if f.endswith(".py"):
                    a = os.path.join(d, f)
                    b = os.path.join(DST, os.path.relpath(a, SRC))
                    assert filecmp.cmp(a, b, shallow=False), b
                    n += 1
    for f in ("main.py", "main_viz.py", "hyperparam.ini"):
        assert filecmp.cmp(os.path.join(SRC, f), os.path.join(DST, f), shallow=False), f
    assert n > 15


def test_reference_imports_from_shipped_copy():
    if not os.path.isdir(DST):
        pytest.skip("baseline/_ref not shipped")
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); from oracle import reference_env as E; d = E.activate(%r); "
            "import disvae, os; assert os.path.realpath(disvae.__file__).startswith(os.path.realpath(d)); "
            "from disvae.training import Trainer; import main; print('ok')" % (ROOT, DST))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
