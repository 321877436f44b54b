"""Glow-curve post-processing on the fused ensemble histograms.

The reference bins one replica's events into 1 degC bins and smooths with a 50 degC boxcar
(``src/class/plots.py:39-47``, ``running_mean`` at ``:19-21``), after averaging ``Lum`` across
replicas BY STEP INDEX on replica 0's time axis (``plots.py:61-62``) -- which mixes events of
different temperatures.  Here every replica's events are binned on the common axis inside the
kernel (integer histogram, ``MCL_AXIS_TEMP``), so the ensemble mean is a proper per-temperature
average; the smoothing is the reference's.
"""
from __future__ import annotations

import numpy as np


def running_mean(a: np.ndarray, k: int = 5) -> np.ndarray:
    """Boxcar mean, ``valid`` mode (reference plots.py:19-21)."""
    return np.convolve(a, np.ones(k) / k, "valid")


def glow_curve(hist_events_row: np.ndarray, n_replicas: int, bin_width: float = 1.0,
               win_deg: float = 50.0) -> np.ndarray:
    """TL intensity per degC per replica, smoothed like ``hist_and_smooth`` (plots.py:39-47).

    ``hist_events_row`` is one row of the kernel's event histogram on a temperature axis whose
    bins are ``bin_width`` degC wide.
    """
    hist = np.asarray(hist_events_row, dtype=np.float64) / float(n_replicas) / bin_width
    k = max(1, int(win_deg / bin_width))
    return running_mean(hist, k=k)


def decay_curve(hist_occ_row: np.ndarray, hist_occ_sq_row: np.ndarray, n_replicas: int, n_traps: float):
    """Ensemble mean and standard deviation of the filled fraction n(t)/N_e at every bin edge."""
    n = float(n_replicas)
    mean = np.asarray(hist_occ_row, dtype=np.float64) / n
    var = np.maximum(np.asarray(hist_occ_sq_row, dtype=np.float64) / n - mean * mean, 0.0)
    return mean / n_traps, np.sqrt(var) / n_traps
