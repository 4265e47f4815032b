import sys, numpy as np
sys.path.insert(0, ".")
import apda_fft_b200
from oracle import c_oracle
an = apda_fft_b200.Analyzer(0)
for log2n in (14, 16):
    x = np.round(np.random.default_rng(1).standard_normal(1 << log2n), 6)
    try:
        got = an.fft(x)
        want = c_oracle.start_fft_batch(x)
        print(log2n, "equal:", np.array_equal(got.view(np.float64), want.view(np.float64)))
    except Exception as e:
        print(log2n, "ERR", e)
        break
