"""ORACLE (test infrastructure, never the product path).

``oracle/_ref/``: a build-time, git-ignored copy of the UNMODIFIED reference checkout (pure Python, 10 files), made
by ``__graft_entry__.build()`` wherever ``/root/reference`` exists.  It is not listed in ``.gpurunignore``, so it
travels to the GPU box with the snapshot like the built ``.so`` files, and gives that box

  * the real ``Gateway.work_flow_fft`` call site (GT_FFT_v5.py:620-680) for the drop-in replay test, and
  * the untouched ``start_fft`` + pickers for the CPU baseline of ``bench.py`` (``kind: "reference"``).

Nothing here is imported by the product; the sources are never committed.
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("APDA_REFERENCE_DIR", "/root/reference")
REF_DIR = os.path.join(_HERE, "_ref")
HOT_PATH = ("metrics/fft_iterativa.py", "utils/get_peak_prominence.py", "utils/get_peak_resolution.py",
            "utils/load_data.py")
CALL_SITE = ("GT_FFT_v5.py", "protocol_decoder.py", "protocol_radio.py", "utils/ftp_manager.py",
             "utils/fastapi_manager.py")


def build_ref(force: bool = False) -> str | None:
    """Copy the reference's .py files into oracle/_ref/ (only where the reference checkout is present)."""
    if not os.path.isdir(REF_SRC):
        return REF_DIR if available() else None
    for root, _dirs, files in os.walk(REF_SRC):
        for name in files:
            if not name.endswith(".py"):
                continue
            src = os.path.join(root, name)
            dst = os.path.join(REF_DIR, os.path.relpath(src, REF_SRC))
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if force or not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read():
                shutil.copyfile(src, dst)
    return REF_DIR


def available(call_site: bool = False) -> bool:
    need = HOT_PATH + (CALL_SITE if call_site else ())
    return all(os.path.exists(os.path.join(REF_DIR, rel)) for rel in need)


class RefModules:
    """The reference's hot-path functions, loaded from oracle/_ref under private module names (so they never collide
    with the drop-in ``metrics`` / ``utils`` packages of apda-fft_b200/)."""

    def __init__(self):
        if not available():
            raise FileNotFoundError("oracle/_ref is not populated (run __graft_entry__.build() where /root/reference exists)")
        fft = _load("_apda_ref_fft_iterativa", "metrics/fft_iterativa.py")
        prom = _load("_apda_ref_get_peak_prominence", "utils/get_peak_prominence.py")
        res = _load("_apda_ref_get_peak_resolution", "utils/get_peak_resolution.py")
        self.start_fft = fft.start_fft
        self.get_top_peaks_prominence = prom.get_top_peaks_prominence
        self.get_top_peaks_resolution = res.get_top_peaks_resolution


def _load(name: str, rel: str):
    mod = sys.modules.get(name)
    if mod is None:
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_DIR, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    return mod
