"""Per-window peak records (include/apda_b200.h: apda_peak_rec) <-> the reference's list-of-dict results.

The device decides WHICH peaks are reported (all comparisons, the rounded-magnitude ordering, the hump
exclusion, the resolution criterion).  The presentation fields the reference derives with Python's
round()/float arithmetic (freq, damping, q-factor) are recomputed here from (idx, width_bins, mag,
prominence, fs, n) in the reference's operation order, so they are the same doubles.
"""
from __future__ import annotations

import numpy as np


def record_dtype(rec_cap: int = 5) -> np.dtype:
    peak = np.dtype([("idx", "<i4"), ("width_bins", "<i4"), ("mag", "<f8"), ("prominence", "<f8")])
    dt = np.dtype([("count", "<i4"), ("status", "<i4"), ("pk", peak, (rec_cap,))])
    assert dt.itemsize == 8 + 24 * rec_cap
    return dt


def prominence_dicts(rec, fs: float, n: int) -> list[dict]:
    """utils/get_peak_prominence.py:187-194 record layout (reference), one window."""
    out = []
    df = fs / n
    for a in range(int(rec["count"])):
        pk = rec["pk"][a]
        idx = int(pk["idx"])
        fn = idx * df
        q_factor = fn / (int(pk["width_bins"]) * df)
        damping = 1 / (2 * q_factor)
        out.append({"freq": round(fn, 4), "mag": round(float(pk["mag"]), 4), "prominence": float(pk["prominence"]),
                    "damping": round(damping * 100, 2), "q-factor": round(q_factor, 2), "idx": idx})
    return out


def resolution_dicts(rec, fs: float, n: int) -> list[dict]:
    """utils/get_peak_resolution.py:113 record layout (reference), one window."""
    df = fs / n
    return [{"freq": int(pk["idx"]) * df, "mag": float(pk["mag"]), "idx": int(pk["idx"])}
            for pk in rec["pk"][: int(rec["count"])]]


def gateway_entry(peaks: list[dict]) -> dict:
    """The per-axis ``fft_dict`` entry the reference's caller builds from a picker result (GT_FFT_v5.py:644-659):
    ``peak_freq`` / ``max_mag`` of the first peak (-1 when there is none) plus ``peak_freq_i`` / ``max_mag_i`` for
    every returned peak.  (The timing fields the caller adds afterwards are its own business.)"""
    entry = {"peak_freq": -1, "max_mag": -1}
    if peaks:
        entry["peak_freq"] = peaks[0]["freq"]
        entry["max_mag"] = peaks[0]["mag"]
        for i, pk in enumerate(peaks):
            entry[f"peak_freq_{i + 1}"] = pk["freq"]
            entry[f"max_mag_{i + 1}"] = pk["mag"]
    return entry


def fleet_table(recs, fs, n: int, flexible: bool = True, k: int = 4):
    """Columnar view of a record batch for bulk consumers: (count[B], idx[B,k], freq[B,k], mag[B,k]).
    freq/mag are the raw doubles (idx * fs / n and |X|); apply prominence_dicts per window where the reference's
    decimal rounding is needed."""
    import numpy as np
    recs = np.asarray(recs)
    idx = recs["pk"]["idx"][:, :k].astype(np.int64)
    fs_col = np.broadcast_to(np.asarray(fs, dtype=np.float64), (recs.shape[0],))[:, None]
    freq = np.where(idx >= 0, idx * (fs_col / n), np.nan)
    mag = np.where(idx >= 0, recs["pk"]["mag"][:, :k], np.nan)
    return recs["count"].astype(np.int64), idx, freq, mag


def upload_metrics(summary: dict, axis: str, fft_entry: dict) -> dict:
    """The ``metriche`` block the reference's uploader derives for one file (utils/fastapi_manager.py:37-47, 58-64):
    RMS-vector angles from the summary line, the axis' own RMS and the top-4 peak arrays of ``gateway_entry``
    (missing peaks are 0.0, as ``current_fft.get(..., 0.0)`` does there).  Pure host arithmetic, same operation order."""
    from math import acos, atan2, degrees
    rms = {"X": summary["rms_x"], "Y": summary["rms_y"], "Z": summary["rms_z"]}
    norm = (rms["X"] ** 2 + rms["Y"] ** 2 + rms["Z"] ** 2) ** 0.5          # length of the RMS vector
    azimuth = degrees(atan2(rms["Y"], rms["X"]))
    inclination = degrees(acos(rms["Z"] / norm)) if norm != 0 else 0
    top4 = range(1, 5)
    return {
        "temp": summary["temperature"],
        "humidity": summary.get("humidity", 0.0),
        "phi": azimuth,
        "theta": inclination,
        "rms_asse": rms.get(axis, 0.0),
        "fft_freqs": [fft_entry.get(f"peak_freq_{i}", 0.0) for i in top4],
        "fft_mags": [fft_entry.get(f"max_mag_{i}", 0.0) for i in top4],
    }


def fleet_arrow(recs, fs, n: int, flexible: bool = True, k: int = 4, first_window: int = 0):
    """Columnar table (pyarrow) of a record batch for bulk consumers on rank 0: one row per window,
    list columns ``idx`` / ``freq`` / ``mag`` hold the ``count`` peaks in the picker's output order (raw doubles)."""
    import numpy as np
    import pyarrow as pa
    count, idx, freq, mag = fleet_table(recs, fs, n, flexible, k)
    count = np.minimum(count, k)
    offsets = np.concatenate([[0], np.cumsum(count)]).astype(np.int32)
    keep = np.arange(k)[None, :] < count[:, None]

    def ragged(values, typ):
        return pa.ListArray.from_arrays(pa.array(offsets), pa.array(values[keep], type=typ))
    return pa.table({"window": pa.array(np.arange(first_window, first_window + count.shape[0], dtype=np.int64)),
                     "count": pa.array(count), "idx": ragged(idx, pa.int64()), "freq": ragged(freq, pa.float64()),
                     "mag": ragged(mag, pa.float64())})


def write_fleet_jsonl(path, recs, fs, n: int, flexible: bool = True, k: int = 4, first_window: int = 0) -> int:
    """One JSON object per window: ``{"window", "fft_freqs", "fft_mags"}`` with the reference's rounded values
    (prominence_dicts / resolution_dicts), padded with 0.0 to k entries like the uploader's payload.  Returns the rows."""
    import json
    import numpy as np
    recs = np.asarray(recs)
    fs_col = np.broadcast_to(np.asarray(fs, dtype=np.float64), (recs.shape[0],))
    conv = prominence_dicts if flexible else resolution_dicts
    with open(path, "w") as fh:
        for w in range(recs.shape[0]):
            peaks = conv(recs[w], float(fs_col[w]), n)[:k]
            row = {"window": first_window + w,
                   "fft_freqs": [p["freq"] for p in peaks] + [0.0] * (k - len(peaks)),
                   "fft_mags": [p["mag"] for p in peaks] + [0.0] * (k - len(peaks))}
            fh.write(json.dumps(row) + "\n")
    return int(recs.shape[0])
