"""Synthetic multi-tone accelerometer windows (host generator).

Two families, both defined in SURVEY.md Appendix B (they are *our* test inputs,
the reference ships no data):

* ``kat_window``   - the fixed known-answer windows KAT-A/B/C (B.1): one LCG
  noise stream seeded with 42, ``math.sin`` tones, 6-decimal quantisation
  (mirrors the ``"%8.6f"`` sample text of the sensor logs,
  reference protocol_decoder.py:174).
* ``fleet_windows`` - the counter-hash fleet generator (B.2): every window
  ``w`` derives its tones, phases and noise seed from splitmix64(seed, w), so
  any rank can generate any shard without communication.

Everything is integer/IEEE-double arithmetic with a fixed operation order, so
the same window index gives the same bits on every host.
"""
from __future__ import annotations

import math

import numpy as np

_M64 = (1 << 64) - 1
_GOLDEN = 0x9E3779B97F4A7C15
_LCG_A = 6364136223846793005
_LCG_C = 1442695040888963407

KAT_TONES = {
    # name: (N, fs, ((cycles, amplitude, phase), ...))
    "A": (1024, 125.0, ((25.4, 0.5, 0.0), (63.1, 0.3, 0.3), (124.5, 0.2, 1.1))),
    "B": (4096, 125.0, ((101.6, 0.5, 0.0), (252.4, 0.3, 0.3), (498.0, 0.2, 1.1))),
    "C": (8192, 250.0, ((203.2, 0.5, 0.0), (504.8, 0.3, 0.3), (996.0, 0.2, 1.1))),
}

KAT0_INPUT = [-21, 16, -11, 26, -1, -28, 9, -18]


def splitmix64(z: int) -> int:
    z = (z + _GOLDEN) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def window_uniforms(w: int, seed: int = 42) -> list[float]:
    """U_0..U_9 in [0,1) for window ``w`` (Appendix B.2)."""
    z0 = (seed * _GOLDEN + w) & _M64
    return [(splitmix64((z0 + k * _GOLDEN) & _M64) >> 11) / 9007199254740992.0 for k in range(10)]


def window_params(w: int, n: int, seed: int = 42, on_bin: bool = False):
    """(cycles[3], amplitudes[3], phases[3], lcg_state) of fleet window ``w``."""
    u = window_uniforms(w, seed)
    c = [(0.025 + 0.010 * u[0]) * n, (0.055 + 0.015 * u[1]) * n, (0.095 + 0.020 * u[2]) * n]
    if on_bin:
        c = [float(round(v)) for v in c]
    a = [0.5 * (0.9 + 0.2 * u[3]), 0.3 * (0.9 + 0.2 * u[4]), 0.2 * (0.9 + 0.2 * u[5])]
    phi = [2.0 * math.pi * u[6], 2.0 * math.pi * u[7], 2.0 * math.pi * u[8]]
    state = int(u[9] * 9007199254740992.0) | 1
    return c, a, phi, state


def _lcg_noise(state: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.float64)
    s = state
    for i in range(n):
        s = (s * _LCG_A + _LCG_C) & _M64
        out[i] = (s >> 11) / 9007199254740992.0 * 2.0 - 1.0
    return out


def _lcg_noise_block(states: np.ndarray, n: int) -> np.ndarray:
    """Vectorised over windows: states is uint64[B]; returns float64[B, n]."""
    s = states.astype(np.uint64).copy()
    out = np.empty((s.shape[0], n), dtype=np.float64)
    a = np.uint64(_LCG_A)
    c = np.uint64(_LCG_C)
    with np.errstate(over="ignore"):
        for i in range(n):
            s = s * a + c
            out[:, i] = (s >> np.uint64(11)).astype(np.float64) / 9007199254740992.0 * 2.0 - 1.0
    return out


def fleet_window(w: int, n: int, seed: int = 42, on_bin: bool = False) -> np.ndarray:
    """One fleet window as float64[n] (numpy sin, then np.round(., 6))."""
    c, a, phi, state = window_params(w, n, seed, on_bin)
    i = np.arange(n, dtype=np.float64)
    x = np.zeros(n, dtype=np.float64)
    for t in range(3):
        x = x + a[t] * np.sin(2.0 * np.pi * c[t] * i / n + phi[t])
    x = x + 0.01 * _lcg_noise(state, n)
    return np.round(x, 6)


def fleet_windows(first: int, count: int, n: int, seed: int = 42, on_bin: bool = False,
                  dtype=np.float64) -> np.ndarray:
    """Windows first..first+count-1 as dtype[count, n]; bit-equal to fleet_window per row."""
    cs = np.empty((count, 3)); am = np.empty((count, 3)); ph = np.empty((count, 3))
    st = np.empty(count, dtype=np.uint64)
    for r in range(count):
        c, a, phi, state = window_params(first + r, n, seed, on_bin)
        cs[r] = c; am[r] = a; ph[r] = phi; st[r] = state
    i = np.arange(n, dtype=np.float64)[None, :]
    x = np.zeros((count, n), dtype=np.float64)
    for t in range(3):
        x = x + am[:, t:t + 1] * np.sin(2.0 * np.pi * cs[:, t:t + 1] * i / n + ph[:, t:t + 1])
    x = x + 0.01 * _lcg_noise_block(st, n)
    return np.round(x, 6).astype(dtype)


def kat_window(name: str):
    """(samples float64[N], fs) of KAT-A/B/C (Appendix B.1; math.sin, Python round)."""
    n, fs, tones = KAT_TONES[name]
    s = 42
    x = np.empty(n, dtype=np.float64)
    for i in range(n):
        s = (s * _LCG_A + _LCG_C) & _M64
        u = (s >> 11) / 9007199254740992.0 * 2.0 - 1.0
        acc = 0.0
        for (c, a, phi) in tones:
            acc += a * math.sin(2.0 * math.pi * c * i / n + phi)
        x[i] = round(acc + 0.01 * u, 6)
    return x, fs


def noise_window(w: int, n: int, seed: int = 7) -> np.ndarray:
    """Noise-only window (many threshold candidates; stresses the general picker path)."""
    state = int(window_uniforms(w, seed)[9] * 9007199254740992.0) | 1
    return np.round(_lcg_noise(state, n), 6)
