"""One small launch of every kernel family (for compute-sanitizer memcheck / racecheck runs on the B200 box)."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import apda_fft_b200
import apda_fft_b200.synth as synth
import cases

an = apda_fft_b200.Analyzer(0)
ok = []
for n in (1024, 4096):
    x = synth.fleet_windows(0, 6, n)
    for dt in (np.float64, np.float32):
        xs = x.astype(dt)
        an.fft(xs); an.analyze(xs, 125.0, flexible=True); an.analyze(xs, 125.0, flexible=False)
        an.analyze(xs[:, : n - 37], 125.0, flexible=True)
        an.ctx.set_generic_only(True)
        an.fft(xs); an.analyze(xs, 125.0, flexible=True); an.analyze(xs, 125.0, flexible=False)
        an.ctx.set_generic_only(False)
    an.analyze_fused(x.astype(np.float32), 125.0, flexible=True)
    an.analyze_fused(x.astype(np.float32), 125.0, flexible=False, center=apda_fft_b200._cabi.CENTER_MEAN)
    noise = np.stack([synth.noise_window(w, n) for w in range(4)]).astype(np.float32)
    an.analyze(noise, 125.0, flexible=False); an.analyze_fused(noise, 125.0, flexible=False)
    ok.append(n)
z = np.round(np.random.default_rng(0).standard_normal(1 << 15), 6)
for dt in (np.float64, np.float32):
    s = an.fft(z.astype(dt)); an.peaks(s, 250.0, flexible=True); an.peaks(s, 250.0, flexible=False)
z = np.round(np.random.default_rng(1).standard_normal(1 << 17), 6)
s = an.fft(z); an.peaks(s, 250.0, flexible=True); an.peaks(s, 250.0, flexible=False)
an.fft_c2c(np.random.default_rng(2).standard_normal(256) + 0j)
pay = np.stack([cases.wire_payload(2, 4096), cases.wire_payload(5, 4096, False)])
an.decode_wire16(pay, [0.5, 0.1]); an.analyze_wire16(pay, [0.5, 0.1], 125.0, dtype="f32"); an.analyze_wire16(pay, [0.5, 0.1], 125.0)
print("sanitize smoke done", ok)
