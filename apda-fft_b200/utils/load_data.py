"""Drop-in for the reference module ``utils/load_data.py``: sensor ``.log`` text -> dict.

File format (reference :29-82): line 0 ``<timestamp>;<range>;<fs> Hz;<A> axis;``, line 1 sync state, line 2 five
floats (temperature, rms x/y/z, humidity), line 3 three floats (first x/y/z), lines 4.. ``;``-separated decimal
samples.  Unparsable and non-finite sample tokens are dropped; fewer than 5 lines -> None.

``load_sensor(filepath)`` keeps the reference's signature and result.  ``load_sensors(filepaths)`` is the batched
ingest: the header lines are parsed here, the sample lines of all files go through one
``apda_parse_samples_f64_host`` call (GPU text parser); a file holding a piece in a float() syntax the kernel does not
decide is re-parsed on the host, so every file yields exactly what ``load_sensor`` would.
"""
from __future__ import annotations

import ctypes
import io
import math

import numpy as np

_SUMMARY_KEYS = ("temperature", "rms_x", "rms_y", "rms_z", "humidity")
_FIRST_KEYS = ("first_x", "first_y", "first_z")


def _finite_floats(tokens):
    for tok in tokens:
        if not tok:
            continue
        try:
            val = float(tok)
        except ValueError:
            continue
        if math.isfinite(val):
            yield val


def _parse_header(rows):
    stamp, rng, rate, axis = rows[0].strip().split(";")[:4]
    sync = rows[1].strip().replace(";", "")
    metadata = {
        "timestamp": stamp,
        "sensitivity": rng.replace(" ", ""),
        "fs": float(rate.replace(" Hz", "")),
        "axis": axis.replace(" axis", "").replace(" ", "_"),
        "sync_type": sync,
        "is_synced": 1.0 if sync in ("Synced", "Synced2") else 0.0,
    }
    summary = {}
    fields = rows[2].strip().split(";")
    for pos, key in enumerate(_SUMMARY_KEYS):
        summary[key] = float(fields[pos])
    fields = rows[3].strip().split(";")
    for pos, key in enumerate(_FIRST_KEYS):
        summary[key] = float(fields[pos])
    return metadata, summary


def _sample_lines(rows):
    samples = []
    for row in rows:
        samples.extend(_finite_floats(row.strip().split(";")))
    return samples


def load_sensor(filepath):
    with open(filepath, "r", encoding="utf-8") as fh:
        rows = fh.readlines()
    if len(rows) < 5:
        return None
    metadata, summary = _parse_header(rows)
    return {"metadata": metadata, "summary": summary, "samples": _sample_lines(rows[4:])}


def split_log(data: bytes):
    """(header rows, sample-region bytes) of one log, or None when it has fewer than 5 lines (universal newlines)."""
    pos, lines = 0, 0
    n = len(data)
    while pos < n and lines < 4:
        cr, lf = data.find(b"\r", pos), data.find(b"\n", pos)
        ends = [e for e in (cr, lf) if e >= 0]
        if not ends:
            pos = n
            break
        e = min(ends)
        pos = e + 2 if data[e:e + 2] == b"\r\n" else e + 1
        lines += 1
    if lines < 4 or pos >= n:      # fewer than 5 lines in total
        return None
    rows = io.StringIO(data[:pos].decode("utf-8"), newline=None).readlines()
    return rows, data[pos:]


def load_sensors(filepaths, device: int = 0, as_arrays: bool = False):
    """Batched load_sensor: one dict (or None) per path, samples parsed on the GPU.  ``as_arrays`` keeps the samples as
    float64 numpy arrays instead of Python lists."""
    from apda_fft_b200 import _cabi
    heads, regions, order = [], [], []
    out = [None] * len(filepaths)
    for i, path in enumerate(filepaths):
        with open(path, "rb") as fh:
            parts = split_log(fh.read())
        if parts is None:
            continue
        heads.append(_parse_header(parts[0]))
        regions.append(parts[1])
        order.append(i)
    if not order:
        return out
    offsets = np.zeros(len(regions) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in regions], out=offsets[1:])
    text = b"".join(regions)
    n_max = max(1, max((len(r) + 1) // 2 for r in regions))       # a sample needs at least one digit and one separator
    samples = np.empty((len(regions), n_max), dtype=np.float64)
    n_valid = np.zeros(len(regions), dtype=np.int32)
    flags = np.zeros(len(regions), dtype=np.int32)
    p = ctypes.c_void_p
    buf = np.frombuffer(text, dtype=np.uint8)
    _cabi.default_context(device).call("apda_parse_samples_f64_host", p(buf.ctypes.data), p(offsets.ctypes.data),
                                       len(regions), n_max, p(samples.ctypes.data), p(n_valid.ctypes.data),
                                       p(flags.ctypes.data))
    for j, i in enumerate(order):
        if flags[j]:
            rows = io.StringIO(regions[j].decode("utf-8"), newline=None).readlines()
            vals = np.asarray(_sample_lines(rows), dtype=np.float64)
        else:
            vals = samples[j, : n_valid[j]].copy()
        out[i] = {"metadata": heads[j][0], "summary": heads[j][1], "samples": vals if as_arrays else vals.tolist()}
    return out
