"""Run the REFERENCE's own test files for the tuple plumbing of the hot path
(tests/nn/flow/test_coupling.py, test_sequential.py, test_inverted.py) and for the truncated normal
marginal (tests/distribution/test_normal.py) with bgflow's classes replaced by this package's mirrors
(SplitFlow, MergeFlow, SwapFlow, CouplingFlow, WrapFlow, SetConstantFlow, SequentialFlow, InverseFlow,
Flow, Transformer, TruncatedNormalDistribution) — CPU, fp32 and fp64, the reference's generic
(non-kernel) transformers.  The test sources are read from the reference checkout at run time and
executed unmodified apart from the import lines; nothing is copied into this repository.  Build
container only (skipped where /root/reference is absent)."""

import os
import re
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "bgflow")), reason="reference checkout not present")

SHIM = '''
import numpy as np
np.infty = np.inf
import bgflow as _ref
import bgflow.nn.flow as _ref_flow
import bgflow_b200 as _bg

_MIRRORED = ("Flow", "SequentialFlow", "InverseFlow", "SplitFlow", "MergeFlow", "SwapFlow", "CouplingFlow", "WrapFlow",
             "SetConstantFlow", "Transformer", "TruncatedNormalDistribution")
for _n in dir(_ref):
    if not _n.startswith("__"):
        globals()[_n] = getattr(_ref, _n)
for _n in _MIRRORED:
    globals()[_n] = getattr(_bg, _n)


class flow:                      # stands in for `from bgflow.nn import flow`
    pass


for _n in dir(_ref_flow):
    if not _n.startswith("__"):
        setattr(flow, _n, getattr(_ref_flow, _n))
for _n in _MIRRORED:
    setattr(flow, _n, getattr(_bg, _n))
'''

CONFTEST = '''
import pytest, torch

@pytest.fixture(params=["cpu"])
def device(request):
    return torch.device(request.param)

@pytest.fixture(params=[torch.float32, torch.float64])
def dtype(request, device):
    return request.param

@pytest.fixture()
def ctx(dtype, device):
    return {"dtype": dtype, "device": device}
'''


def _rewrite(src):
    src = re.sub(r"^from bgflow\.nn\.flow\.sequential import", "from bgx_mirror_shim import", src, flags=re.M)
    src = re.sub(r"^from bgflow\.nn import flow$", "from bgx_mirror_shim import flow", src, flags=re.M)
    src = re.sub(r"^from bgflow import", "from bgx_mirror_shim import", src, flags=re.M)
    src = re.sub(r"^from bgflow\.distribution import", "from bgx_mirror_shim import", src, flags=re.M)
    return src


@pytest.mark.parametrize("name,select,expected", [
    ("nn/flow/test_coupling.py", None, 18), ("nn/flow/test_sequential.py", None, 2), ("nn/flow/test_inverted.py", None, 10),
    ("distribution/test_normal.py", "truncated_normal", 16)])
def test_reference_plumbing_tests_pass_on_the_mirror(tmp_path, name, select, expected):
    (tmp_path / "bgx_mirror_shim.py").write_text(SHIM)
    (tmp_path / "conftest.py").write_text(CONFTEST)
    src = open(os.path.join(REF, "tests", name)).read()
    name = os.path.basename(name)
    out = _rewrite(src)
    assert "bgx_mirror_shim" in out
    (tmp_path / name).write_text(out)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(tmp_path), ROOT, os.path.join(ROOT, "oracle", "_stubs"), REF])
    sel = ["-k", select] if select else []
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", *sel, str(tmp_path / name)],
                         capture_output=True, text=True, env=env, cwd=str(tmp_path))
    tail = textwrap.shorten(res.stdout[-1500:], 1500)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    m = re.search(r"(\d+) passed", res.stdout)
    assert m and int(m.group(1)) >= expected, tail
