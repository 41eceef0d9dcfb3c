"""CPU oracle for the bgflow coupling-flow hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU (torch-on-CPU, fp32 or fp64)
restatement of the reference's algorithm for the hot path named in
BASELINE.json (coupling blocks: conditioner MLP -> affine / rational-quadratic
spline transform -> log|det J|; Z-matrix <-> Cartesian internal coordinates).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product (``bgflow_b200``) never
imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * plumbing, DenseNet, affine transformer, internal coordinates: pinned against
    outputs of the reference itself (imported from /root/reference in the build
    container; fixtures + generator script under tests/golden/).
  * rational-quadratic spline arithmetic: the reference delegates it to the
    third-party ``nflows`` package (unpinned, not vendored, not installable
    here).  The restatement in ``oracle/flows.py::rational_quadratic_spline`` is
    pinned against an independent on-disk port of the same nflows function
    (transformers' VITS ``_rational_quadratic_spline``), against fp64
    autograd derivatives and against the reference's own property tests.  For
    the ``enable_identity_init=True`` softplus variant there is no third-party
    known-answer vector: that one constant (beta) is "parity unpinned".
"""

from . import flows, ic, cdf  # noqa: F401
