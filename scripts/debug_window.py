import sys
import numpy as np
sys.path.insert(0, ".")
import apda_fft_b200
from apda_fft_b200.records import prominence_dicts
w = int(sys.argv[1]); n = 4096
an = apda_fft_b200.Analyzer(0)
x = apda_fft_b200.synth.fleet_window(w, n)[None, :].astype(np.float32)
for mode in ("fast", "generic"):
    an.ctx.set_generic_only(mode == "generic")
    rec = an.analyze(x, 125.0, flexible=True)[0]
    spec = an.fft(x)[0]
    print(mode, [(p["idx"], p["mag"], p["damping"]) for p in prominence_dicts(rec, 125.0, n)])
    m = np.abs(spec[:2048]).astype(np.float32)
    print("   mags around 137:", m[133:142])
an.ctx.set_generic_only(False)
spec = an.fft(x)
rec2 = an.peaks(spec, 125.0, flexible=True)[0]
print("fast peaks on fast spectrum via host peaks:", [(p["idx"], p["mag"]) for p in prominence_dicts(rec2, 125.0, n)])
m = np.abs(spec[0][:2048].astype(np.complex128))
print("mean", m.mean(), "sd", m.std(ddof=1), "thr", m.mean() + 2 * m.std(ddof=1))
