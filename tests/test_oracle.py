"""T0: the oracle (scalar port + C FFT) replays the committed goldens that were produced by the live reference.
CPU only.  When /root/reference is present (build container) the fixture itself is re-derived and compared."""
import os
import statistics
import subprocess
import sys

import numpy as np
import pytest

import cases
from oracle import c_oracle, ref_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(fn, *args):
    try:
        return {"ok": fn(*args)}
    except Exception as exc:
        return {"error": type(exc).__name__, "message": str(exc)}


@pytest.mark.parametrize("spec", cases.CASES, ids=[c["id"] for c in cases.CASES])
def test_port_and_c_oracle_match_golden(spec, golden):
    g = golden["cases"][spec["id"]]
    x, fs = cases.build_samples(spec)
    assert cases.sha16(x) == g["input_sha"]
    port = ref_port.start_fft(x.tolist(), fs)
    assert len(port) == g["n_fft"]
    assert cases.spectrum_sha16(port) == g["spectrum_sha"]
    assert port[0] == 0 and type(port[0]) is int
    c_spec = c_oracle.start_fft_batch(x)[0]
    assert cases.spectrum_sha16(c_spec) == g["spectrum_sha"]
    for i, (re, im) in g["bins"].items():
        assert complex(port[int(i)]) == complex(re, im)
    assert _run(ref_port.top_peaks_prominence, port, fs) == g["prominence"]
    assert _run(ref_port.top_peaks_resolution, port, fs) == g["resolution"]
    if g["n_fft"] >= 4:
        assert _run(c_oracle.peaks_prominence, c_spec, fs) == g["prominence"]
        assert _run(c_oracle.peaks_resolution, c_spec, fs) == g["resolution"]


def test_survey_known_answers(golden):
    """SURVEY.md Appendix B.1 digests (independently produced during the survey from the live reference)."""
    expect = {"katA": ("53d00f25f6dc7ef6", "8c5d9a4667584d98", [25, 63, 124]),
              "katB": ("edfc2022f3cda128", "6fc481ac7e7bcece", [102, 252, 498]),
              "katC": ("b81ccf581724a8bb", "94d3bcbf65436df0", [203, 505, 996])}
    for cid, (in_sha, sp_sha, idx) in expect.items():
        g = golden["cases"][cid]
        assert (g["input_sha"], g["spectrum_sha"]) == (in_sha, sp_sha)
        assert [p["idx"] for p in g["prominence"]["ok"]] == idx
        assert [p["idx"] for p in g["resolution"]["ok"]] == idx
    katb = golden["cases"]["katB"]["prominence"]["ok"][0]
    assert (katb["freq"], katb["mag"], katb["prominence"], katb["damping"], katb["q-factor"]) == \
        (3.1128, 772.0513, 772.0478456461883, 0.98, 51.0)
    assert ref_port.start_fft([-21, 16, -11, 26, -1, -28, 9, -18], 1.0)[1] == (-19.999999999999993 - 42.22539674441618j)
    assert [p["idx"] for p in golden["cases"]["fleet4096_w999999"]["resolution"]["ok"]] == [127, 284, 450, 123, 131]


@pytest.mark.parametrize("spec", cases.SPECTRA, ids=[c["id"] for c in cases.SPECTRA])
def test_picker_only_spectra(spec, golden):
    g = golden["spectra"][spec["id"]]
    z, fs = cases.build_spectrum(spec)
    for k in (4, 5, 12):
        assert _run(ref_port.top_peaks_prominence, z.tolist(), fs, k) == g[f"prominence_k{k}"]
        assert _run(ref_port.top_peaks_resolution, z.tolist(), fs, k) == g[f"resolution_k{k}"]


def test_k_variants(golden):
    by_id = {c["id"]: c for c in cases.CASES}
    for row in golden["k_variants"]:
        x, fs = cases.build_samples(by_id[row["case"]])
        spec = c_oracle.start_fft_batch(x)[0]
        assert _run(c_oracle.peaks_prominence, spec, fs, row["k"]) == row["prominence"]
        assert _run(c_oracle.peaks_resolution, spec, fs, row["k"]) == row["resolution"]


def test_helper_goldens(golden):
    for blk in golden["helpers"]:
        mags = cases.mags_case(blk["seed"], blk["n"], blk["style"]).tolist()
        for row in blk["rows"]:
            j = row["j"]
            assert ref_port.prominence_of(mags, j) == row["prominence"]
            assert ref_port.half_power_width(mags, row["prominence"], j, blk["fs"], blk["n_fft"]) == row["width_hz"]
            assert ref_port.half_height_bins(mags, j) == row["whm"]
            assert ref_port.resolution_between(mags, j, row["other"]) == row["rs"]


def test_pad_and_empty(golden):
    for n, padded in golden["pad"].items():
        assert len(ref_port.pad_to_pow2([1.5] * int(n))) == padded
        assert c_oracle.padded_len(int(n)) == max(padded, 1)
    assert ref_port.start_fft([], 1.0) == [0]
    assert ref_port.center_on_median([]) == []
    with pytest.raises(statistics.StatisticsError):
        ref_port.top_peaks_prominence([0, 1 + 1j], 1.0)


def test_c_oracle_batch_equals_rowwise():
    import apda_fft_b200.synth as synth
    x = synth.fleet_windows(100, 8, 1024)
    spec = c_oracle.start_fft_batch(x)
    for r in range(8):
        assert np.array_equal(spec[r].view(np.float64), c_oracle.start_fft_batch(x[r])[0].view(np.float64))
        assert cases.spectrum_sha16(ref_port.start_fft(x[r].tolist(), 125.0)) == cases.spectrum_sha16(spec[r])
    z = np.round(np.random.default_rng(0).standard_normal(64) + 1j * np.random.default_rng(0).standard_normal(64), 6)
    assert cases.spectrum_sha16(c_oracle.fft_c2c(z)) == cases.spectrum_sha16(ref_port.dit_radix2(z.tolist()))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference only exists in the build container")
def test_golden_file_still_matches_live_reference():
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "tests", "golden", "make_golden.py"), "--check"],
                         stdout=subprocess.DEVNULL)
    assert rc == 0


def test_wire_decode_port_matches_golden(golden):
    for wc in cases.WIRE_CASES:
        g = golden["wire"][wc["id"]]
        pay = cases.wire_payload(wc["seed"], wc["n"], wc.get("specials", True))
        loaded = ref_port.wire_samples_as_loaded(pay.tolist(), wc["first_value"])
        assert len(loaded) == g["n_valid"] and cases.sha16(np.asarray(loaded)) == g["sha"]
        assert loaded[:6] == g["head"]
        assert ref_port.decode_wire_samples_text(pay.tolist(), wc["first_value"])[:6] == g["text_head"]


def test_log_parsing_port_and_split(golden, tmp_path):
    import io
    sys.path.insert(0, os.path.join(ROOT, "apda-fft_b200"))
    from utils.load_data import load_sensor, split_log
    for lc in cases.LOG_CASES:
        g = golden["logs"][lc["id"]]
        text = cases.log_text(lc["seed"], lc["n"], lc["nasty"], lc.get("newline", "\n"), lc.get("per_line", 60))
        lines = io.StringIO(text, newline=None).readlines()
        got = ref_port.parse_log_sample_lines(lines[4:])
        assert len(got) == g["n_samples"] and cases.sha16(np.asarray(got)) == g["sha"] and got[-4:] == g["tail"]
        path = tmp_path / (lc["id"] + ".log")
        path.write_bytes(text.encode("utf-8"))
        mine = load_sensor(str(path))
        assert mine["samples"] == got and mine["metadata"]["fs"] == g["fs"] and mine["metadata"]["axis"] == g["axis"]
        parts = split_log(text.encode("utf-8"))
        if lc["n"] == 0:
            assert parts is None or ref_port.parse_log_sample_lines(
                io.StringIO(parts[1].decode(), newline=None).readlines()) == []
        else:
            rows = io.StringIO(parts[1].decode("utf-8"), newline=None).readlines()
            assert parts[0] == lines[:4] and ref_port.parse_log_sample_lines(rows) == got


def _fuzz_inputs(seed, count):
    """Deterministic adversarial inputs: ties, plateaus, odd lengths, DC offsets, spiky magnitude spectra."""
    rng = np.random.default_rng(seed)
    for i in range(count):
        n = int(rng.integers(1, 70)) if i % 3 else int(rng.choice([2, 3, 4, 7, 8, 16, 31, 32, 64, 100, 128, 200, 256]))
        kind = i % 4
        if kind == 0:
            x = np.round(rng.standard_normal(n), 1)                       # many ties
        elif kind == 1:
            x = np.round(np.sin(np.arange(n) * rng.uniform(0.1, 2.5)) + 0.05 * rng.standard_normal(n) + rng.uniform(-3, 3), 6)
        elif kind == 2:
            x = rng.integers(-3, 4, n).astype(np.float64)                 # plateaus, repeated medians
        else:
            x = np.round(np.exp(rng.standard_normal(n)), 3)               # skewed
        yield x.tolist()


def _fuzz_spectra(seed, count):
    rng = np.random.default_rng(seed)
    for i in range(count):
        n = int(rng.choice([8, 16, 32, 64, 128, 256]))
        mags = np.round(np.exp(1.4 * rng.standard_normal(n)), int(rng.integers(0, 4)))   # rounded: equal neighbours occur
        mags[0] = 0.0
        phase = np.exp(1j * rng.uniform(0, 2 * np.pi, n)) if i % 2 else np.ones(n)
        yield (mags * phase).tolist(), float(rng.choice([31.25, 62.5, 125.0, 250.0, 500.0])), int(rng.integers(1, 8))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference only exists in the build container")
def test_port_fuzz_against_live_reference():
    """Oracle pin beyond the fixed goldens: 400 adversarial sample windows and 300 hand-shaped spectra through the
    UNMODIFIED reference and through the port - equal objects (spectra bit for bit, peak dicts, exception types)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    ref = make_golden.import_reference()
    for x in _fuzz_inputs(20260101, 400):
        want = _run(ref["fft_iterativa"].start_fft, list(x), 125.0)
        got = _run(ref_port.start_fft, list(x), 125.0)
        assert want.keys() == got.keys(), x
        if "ok" in want:
            assert len(want["ok"]) == len(got["ok"])
            assert all(complex(a) == complex(b) for a, b in zip(want["ok"], got["ok"])), x
            spec = want["ok"]
            for fn_ref, fn_port in ((ref["get_peak_prominence"].get_top_peaks_prominence, ref_port.top_peaks_prominence),
                                    (ref["get_peak_resolution"].get_top_peaks_resolution, ref_port.top_peaks_resolution)):
                assert _run(fn_ref, list(spec), 125.0) == _run(fn_port, list(spec), 125.0), x
            if len(x) >= 2:
                c = c_oracle.start_fft_batch(np.asarray(x, dtype=np.float64))[0]
                assert cases.spectrum_sha16(c) == cases.spectrum_sha16(spec)
    for spec, fs, k in _fuzz_spectra(7, 300):
        assert _run(ref["get_peak_prominence"].get_top_peaks_prominence, list(spec), fs, k) == \
            _run(ref_port.top_peaks_prominence, list(spec), fs, k)
        assert _run(ref["get_peak_resolution"].get_top_peaks_resolution, list(spec), fs, k) == \
            _run(ref_port.top_peaks_resolution, list(spec), fs, k)


def test_reference_copy_and_gateway_replay_reference_mode(golden, tmp_path):
    """oracle/_ref (build-time copy of the unmodified reference; only where /root/reference exists): the hot-path
    functions loaded from it reproduce the goldens, and the real call site Gateway.work_flow_fft runs on a KAT-A log with
    digidevice stubbed - the CPU half of tests/test_gpu_round2.py's drop-in replay."""
    import json
    import subprocess
    import sys
    from oracle import ref_copy
    if ref_copy.build_ref() is None or not ref_copy.available(call_site=True):
        pytest.skip("no reference checkout and no oracle/_ref copy")
    ref = ref_copy.RefModules()
    g = golden["cases"]["katA"]
    x, fs = cases.build_samples(g["spec"])
    spec = ref.start_fft(x.tolist(), fs)
    assert cases.spectrum_sha16(spec) == g["spectrum_sha"]
    assert ref.get_top_peaks_prominence(spec, fs) == g["prominence"]["ok"]
    assert ref.get_top_peaks_resolution(spec, fs) == g["resolution"]["ok"]
    path = tmp_path / "0013a2_Xaxis.log"
    cases.write_sensor_log(path, x, fs, "X", missing_marker_at=9)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for flexible, key in ((1, "prominence"), (0, "resolution")):
        out = subprocess.run([sys.executable, os.path.join(root, "tests", "gateway_replay.py"), "ref", str(flexible), "0013a2",
                              str(path)], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        entry = json.loads(out.stdout.strip().splitlines()[-1])["fft_dict"]["0013a2"]["X"]
        want = g[key]["ok"]
        assert entry["peak_freq"] == want[0]["freq"] and entry["max_mag"] == want[0]["mag"]
        assert [entry[f"peak_freq_{i + 1}"] for i in range(len(want))] == [p["freq"] for p in want]
