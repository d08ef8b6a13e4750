"""Small helpers mirrored from /root/reference/src/lightkurve/utils.py (the parts the hot path uses)."""
import numpy as np

__all__ = ["LightkurveWarning", "LightkurveError", "validate_method", "running_mean"]


class LightkurveWarning(Warning):
    """Class for all Lightkurve warnings (utils.py:547)."""


class LightkurveError(Exception):
    """Class for Lightkurve exceptions."""


def validate_method(method, supported_methods):
    """Raises a ValueError if a method is not supported (utils.py:577-598)."""
    method = method.lower()
    if method in supported_methods:
        return method
    raise ValueError(
        "method '{}' is not supported; "
        "must be one of {}".format(method, supported_methods)
    )


def running_mean(data, window_size):
    """Running mean with a top-hat window (utils.py:374-387)."""
    if window_size > len(data):
        window_size = len(data)
    cumsum = np.cumsum(np.insert(data, 0, 0))
    return (cumsum[window_size:] - cumsum[:-window_size]) / float(window_size)
