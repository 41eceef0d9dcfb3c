"""Stub of the third-party ``nflows`` package (TEST INFRASTRUCTURE).

bgflow's ConditionalSplineTransformer imports
``nflows.transforms.splines.rational_quadratic_spline`` and
``nflows.transforms.base.InputOutsideDomain`` (bgflow/nn/flow/transformer/spline.py:75,
129-130).  nflows is neither vendored in the reference nor installable here, so when
the *reference itself* is imported in the build container (golden generation only)
this stub routes those two names to the oracle's restatement.
"""
