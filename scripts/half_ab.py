#!/usr/bin/env python
"""Same-box A/B of library builds on the half-spectrum pipeline (apda_analyze_f32_dev with its own workspace):
    python scripts/half_ab.py lib_a.so lib_b.so [--windows 800000] [--n 4096]
Each library runs in its own process (APDA_LIB); prints ns per window and the SHA-256 of the record table."""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(windows, n):
    import torch
    import apda_fft_b200
    dev = torch.device("cuda:0")
    an = apda_fft_b200.Analyzer(0)
    stream = torch.cuda.current_stream(dev)
    an.use_stream(stream.cuda_stream)
    x = torch.empty((windows, n), dtype=torch.float32, device=dev)
    rec = torch.zeros((windows, 128), dtype=torch.uint8, device=dev)
    an.synth_device(0, windows, n, "f32", x.data_ptr())
    fn = lambda: an.analyze_device(x.data_ptr(), windows, n, n, "f32", 125.0, rec.data_ptr(), flexible=True, k=4)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = []
    for _ in range(4):
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(5):
            fn()
        z.record(stream)
        torch.cuda.synchronize()
        best.append(a.elapsed_time(z) / 5)
    sha = hashlib.sha256(rec.cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"HALFAB {os.path.basename(os.environ.get('APDA_LIB', 'default'))} n={n} {min(best) * 1e6 / windows:.3f} ns/window rec_sha={sha}", flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a.endswith(".so")]
    windows = int(sys.argv[sys.argv.index("--windows") + 1]) if "--windows" in sys.argv else 800_000
    n = int(sys.argv[sys.argv.index("--n") + 1]) if "--n" in sys.argv else 4096
    if "--child" in sys.argv:
        child(windows, n)
    else:
        for _ in range(2):
            for lib in args:
                env = dict(os.environ, APDA_LIB=os.path.abspath(lib))
                out = subprocess.run([sys.executable, __file__, "--child", "--windows", str(windows), "--n", str(n)],
                                     env=env, capture_output=True, text=True)
                print("\n".join(l for l in out.stdout.splitlines() if l.startswith("HALFAB")) or out.stderr[-400:], flush=True)
