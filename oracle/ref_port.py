"""ORACLE (test infrastructure, never the product path).

Scalar CPU restatement of the reference's spectral hot path, written from the
behaviour of the reference (file:line cited per function; paths relative to the
reference checkout).  It is pinned against the live reference by
``tests/golden/make_golden.py`` (run in the build container, where the
reference is importable) and the committed fixtures under ``tests/golden/``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module.  The product (``apda-fft_b200``) never does.

The arithmetic is Python ``float``/``complex`` throughout, i.e. IEEE binary64
with separate roundings (no FMA), exactly the type the reference computes in.
"""
from __future__ import annotations

import cmath
import statistics

# ----------------------------------------------------------------------------
# FFT front end  (reference: metrics/fft_iterativa.py)
# ----------------------------------------------------------------------------


def center_on_median(samples):
    """metrics/fft_iterativa.py:5-11 - subtract statistics.median; empty input is returned as is."""
    if len(samples) == 0:
        return samples
    mid = statistics.median(samples)
    return [v - mid for v in samples]


def pad_to_pow2(seq):
    """metrics/fft_iterativa.py:13-22 - right-pad with integer zeros up to the next power of two (len 0 -> 1)."""
    target = 1
    while target < len(seq):
        target *= 2
    return list(seq) + [0] * (target - len(seq))


def stage_twiddles(half):
    """Twiddles of the stage whose butterflies span ``half`` (m = 2*half).

    metrics/fft_iterativa.py:53,57,68 - w_m = exp(-2j*pi/m); w starts at 1+0j for
    every block and is advanced by ``w *= w_m`` after each butterfly, so the
    j-th butterfly of every block of a stage sees the same value: a table.
    """
    step = cmath.exp(-2.0j * cmath.pi / (2 * half))
    tab = []
    w = 1.0 + 0j
    for _ in range(half):
        tab.append(w)
        w *= step
    return tab


def bitrev_indices(n):
    """Index map of metrics/fft_iterativa.py:24-36 (in-place bit-reversal permutation of a length-2^k list)."""
    bits = n.bit_length() - 1
    rev = [0] * n
    for i in range(1, n):
        rev[i] = (rev[i >> 1] >> 1) | ((i & 1) << (bits - 1)) if bits else 0
    return rev


def dit_radix2(x):
    """metrics/fft_iterativa.py:38-70 - forward, unscaled, decimation-in-time radix-2 FFT of a length-2^k list.

    Butterfly (lines 61-65): v = x[hi]*w ; x[lo] = u+v ; x[hi] = u-v, with Python's
    complex product (ac-bd, ad+bc).  Returns a new list (the reference permutes and
    overwrites its argument; callers here never rely on that aliasing).
    """
    n = len(x)
    rev = bitrev_indices(n)
    y = [x[rev[i]] for i in range(n)]
    half = 1
    while 2 * half <= n:
        tab = stage_twiddles(half)
        span = 2 * half
        for base in range(0, n, span):
            for j in range(half):
                lo = base + j
                hi = lo + half
                u = y[lo]
                v = y[hi] * tab[j]
                y[lo] = u + v
                y[hi] = u - v
        half = span
    return y


def start_fft(samples, fs):
    """metrics/fft_iterativa.py:74-87 - centre, pad, transform, then force bin 0 to integer 0.  ``fs`` is unused."""
    spectrum = dit_radix2(pad_to_pow2(center_on_median(samples)))
    spectrum[0] = 0
    return spectrum


# ----------------------------------------------------------------------------
# Shared picker front end
# ----------------------------------------------------------------------------


def half_magnitudes(spectrum):
    """utils/get_peak_prominence.py:159 / utils/get_peak_resolution.py:84 - abs() of bins [0, len//2)."""
    return [abs(spectrum[i]) for i in range(len(spectrum) // 2)]


def noise_threshold(mags):
    """utils/get_peak_prominence.py:163-165 / get_peak_resolution.py:88-90 - (mean, sample stdev, mean+2*stdev)."""
    mu = statistics.mean(mags)
    sd = statistics.stdev(mags)
    return mu, sd, mu + 2 * sd


# ----------------------------------------------------------------------------
# Flexible-structure picker  (reference: utils/get_peak_prominence.py)
# ----------------------------------------------------------------------------


def prominence_of(mags, j):
    """utils/get_peak_prominence.py:32-54 - height above the higher of the two flanking valleys.

    Each side is walked until the first bin strictly higher than the peak (or the array end).
    """
    top = mags[j]
    floor_l = top
    i = j - 1
    while i >= 0 and not mags[i] > top:
        if mags[i] < floor_l:
            floor_l = mags[i]
        i -= 1
    floor_r = top
    i = j + 1
    while i < len(mags) and not mags[i] > top:
        if mags[i] < floor_r:
            floor_r = mags[i]
        i += 1
    return top - max(floor_l, floor_r)


def half_power_bins(mags, prom, j):
    """Bin count of utils/get_peak_prominence.py:89-112 (the -3 dB-of-prominence width, before the *fs/n scaling)."""
    top = mags[j]
    level = (top - prom) + (prom * 0.707)
    lo = j
    while lo > 0 and mags[lo] > level:
        if mags[lo] > top:
            break
        lo -= 1
    hi = j
    while hi < len(mags) - 1 and mags[hi] > level:
        if mags[hi] > top:
            break
        hi += 1
    return max(hi - lo, 1)


def half_power_width(mags, prom, j, fs, n):
    """utils/get_peak_prominence.py:89-112 - width in Hz."""
    return half_power_bins(mags, prom, j) * (fs / n)


def top_peaks_prominence(spectrum, fs, k=4):
    """utils/get_peak_prominence.py:149-226."""
    n = len(spectrum)
    half = n // 2
    mags = half_magnitudes(spectrum)
    df = fs / n
    _, sd, thr = noise_threshold(mags)

    found = []
    for j in range(1, half - 1):
        m = mags[j]
        if not (m > mags[j - 1] and m > mags[j + 1] and m > thr):
            continue
        prom = prominence_of(mags, j)
        if not prom > 0.5 * sd:
            continue
        width = half_power_width(mags, prom, j, fs, n)
        if not width > 0:
            continue
        fn = j * df
        q = fn / width
        damping = 1 / (2 * q)
        if 0.001 <= damping <= 0.07:
            found.append({"freq": round(fn, 4), "mag": round(m, 4), "prominence": prom,
                          "damping": round(damping * 100, 2), "q-factor": round(q, 2), "idx": j})

    found.sort(key=lambda p: p["mag"], reverse=True)      # stable: ties keep ascending idx

    kept = []
    for cand in found:
        hump = False
        for acc in kept:
            if abs(cand["freq"] - acc["freq"]) / acc["freq"] < 0.05:
                if cand["prominence"] / cand["mag"] < 0.10:
                    hump = True
                    break
        if not hump:
            kept.append(cand)
        if len(kept) >= k:
            break
    return kept


# ----------------------------------------------------------------------------
# Rigid-structure picker  (reference: utils/get_peak_resolution.py)
# ----------------------------------------------------------------------------


def half_height_bins(mags, j):
    """utils/get_peak_resolution.py:30-44 - run length above 0.707*mag[j] (right bound may reach len)."""
    level = 0.707 * mags[j]
    lo = j
    while lo > 0 and mags[lo] > level:
        lo -= 1
    hi = j
    while hi < len(mags) and mags[hi] > level:
        hi += 1
    return hi - lo


def resolution_between(mags, a, b):
    """utils/get_peak_resolution.py:48-62."""
    wsum = half_height_bins(mags, a) + half_height_bins(mags, b)
    if wsum == 0:
        return 0
    return 1.18 * abs(b - a) / wsum


def top_peaks_resolution(spectrum, fs, k=5):
    """utils/get_peak_resolution.py:80-128."""
    n = len(spectrum)
    half = n // 2
    mags = half_magnitudes(spectrum)
    df = fs / n
    _, _, thr = noise_threshold(mags)
    freqs2_minus_1 = 2 * df - 1 * df if half > 2 else None     # frequencies[2]-frequencies[1] (:116)

    peaks = []
    while len(peaks) < k:
        best = -1
        best_j = -1
        for j in range(1, half - 1):
            m = mags[j]
            if m > mags[j - 1] and m > mags[j + 1] and m > best and m > thr:
                best = m
                best_j = j
        if best_j < 0:
            break
        f = best_j * df
        if all(resolution_between(mags, p["idx"], best_j) >= 1.5 for p in peaks):
            peaks.append({"freq": f, "mag": best, "idx": best_j})
        reach = round((f * 0.02) / freqs2_minus_1)
        for j in range(max(0, best_j - reach), min(half, best_j + reach + 1)):
            mags[j] = 0
    return peaks


# ----------------------------------------------------------------------------
# Wire-format samples  (reference: protocol_decoder.py, utils/load_data.py)
# ----------------------------------------------------------------------------


def decode_wire_float16(high_byte, low_byte):
    """protocol_decoder.py:116-144 - the sensors' 16-bit sample: 1 sign, 5 exponent, 10 mantissa bits.

    Exponent 31 -> inf (mantissa 0, always positive) or nan; exponent 0 -> sign * 0.00006103515 * (m/1024)
    (a non-IEEE subnormal scale; +0.0 when the mantissa is 0); otherwise sign * 2^(e-15) * (1 + m/1024).
    """
    word = (high_byte << 8) | low_byte
    expo = (word & 0x7C00) >> 10
    sign = -1 if word & 0x8000 else 1
    frac = (word & 0x03FF) / 1024.0
    if expo == 31:
        return float("nan") if frac != 0 else float("inf")
    if expo == 0:
        return sign * 0.00006103515 * frac if frac != 0 else 0.0
    return sign * (pow(2, expo - 15) * (1.0 + frac))


def decode_wire_samples_text(raw_payload, first_value=0.0):
    """protocol_decoder.py:146-175 - byte pairs -> '%8.6f' strings of value + first_value (a trailing odd byte is ignored)."""
    out = []
    for i in range(0, len(raw_payload) - 1, 2):
        out.append("%8.6f" % (decode_wire_float16(raw_payload[i], raw_payload[i + 1]) + first_value))
    return out


def wire_samples_as_loaded(raw_payload, first_value=0.0):
    """What the FFT finally sees: the text above parsed back by utils/load_data.py:67-80 (float(); non-finite dropped)."""
    import math
    vals = []
    for tok in decode_wire_samples_text(raw_payload, first_value):
        v = float(tok)
        if math.isfinite(v):
            vals.append(v)
    return vals


# ----------------------------------------------------------------------------
# Sensor log sample lines  (reference: utils/load_data.py)
# ----------------------------------------------------------------------------


def parse_log_sample_lines(lines):
    """utils/load_data.py:67-80 - lines[4:] of a sensor log: ';'-separated pieces, float(), finite values only."""
    import math
    out = []
    for line in lines:
        for piece in line.strip().split(";"):
            if not piece:
                continue
            try:
                val = float(piece)
            except ValueError:
                continue
            if math.isfinite(val):
                out.append(val)
    return out
