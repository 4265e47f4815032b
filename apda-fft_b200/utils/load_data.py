"""Drop-in for the reference module ``utils/load_data.py``: sensor ``.log`` text -> dict.

File format (reference :29-82): line 0 ``<timestamp>;<range>;<fs> Hz;<A> axis;``, line 1 sync state, line 2 five
floats (temperature, rms x/y/z, humidity), line 3 three floats (first x/y/z), lines 4.. ``;``-separated decimal
samples.  Unparsable and non-finite sample tokens are dropped; fewer than 5 lines -> None.
Host-side text parsing is not part of the GPU path (a batched parser is listed as a next step in DESIGN.md).
"""
from __future__ import annotations

import math

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


def load_sensor(filepath):
    with open(filepath, "r", encoding="utf-8") as fh:
        rows = fh.readlines()
    if len(rows) < 5:
        return None

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

    samples = []
    for row in rows[4:]:
        samples.extend(_finite_floats(row.strip().split(";")))
    return {"metadata": metadata, "summary": summary, "samples": samples}
