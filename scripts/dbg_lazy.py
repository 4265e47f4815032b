import sys, numpy as np
sys.path.insert(0, '/root/repo')
import apda_fft_b200
an = apda_fft_b200.Analyzer(0)
rng = np.random.default_rng(77)
n, b = 8192, 64
z = (rng.standard_normal((b, n)) + 1j * rng.standard_normal((b, n))).astype(np.complex64)
z[:, 0] = 0
fast = an.peaks(z, 125.0, flexible=True)
an.ctx.set_generic_only(True)
slow = an.peaks(z, 125.0, flexible=True)
an.ctx.set_generic_only(False)
bad = [w for w in range(b) if fast[w].tobytes() != slow[w].tobytes()]
print(len(bad), bad[:10])
for w in bad[:3]:
    print("fast", fast[w]); print("slow", slow[w])
