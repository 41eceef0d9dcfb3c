import torch

from oracle.flows import rational_quadratic_spline as _rqs
from .base import InputOutsideDomain

DEFAULT_MIN_BIN_WIDTH = 1e-3
DEFAULT_MIN_BIN_HEIGHT = 1e-3
DEFAULT_MIN_DERIVATIVE = 1e-3


def rational_quadratic_spline(inputs, unnormalized_widths, unnormalized_heights,
                              unnormalized_derivatives, inverse=False, left=0.0, right=1.0,
                              bottom=0.0, top=1.0, min_bin_width=DEFAULT_MIN_BIN_WIDTH,
                              min_bin_height=DEFAULT_MIN_BIN_HEIGHT,
                              min_derivative=DEFAULT_MIN_DERIVATIVE, enable_identity_init=False):
    if torch.min(inputs) < left or torch.max(inputs) > right:
        raise InputOutsideDomain()
    return _rqs(inputs, unnormalized_widths, unnormalized_heights, unnormalized_derivatives,
                inverse=inverse, left=left, right=right, bottom=bottom, top=top,
                min_bin_width=min_bin_width, min_bin_height=min_bin_height,
                min_derivative=min_derivative, enable_identity_init=enable_identity_init)
