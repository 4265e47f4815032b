"""Debug helper (GPU): run the fp32 fleet pipeline over W windows and report windows whose flexible-picker result
differs from the fp64 pipeline on the same samples."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import apda_fft_b200
from apda_fft_b200.records import record_dtype, prominence_dicts

W = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n = 4096
an = apda_fft_b200.Analyzer(0)
dev = torch.device("cuda:0")
an.use_stream(torch.cuda.current_stream(dev).cuda_stream)
d_x = torch.empty((W, n), dtype=torch.float32, device=dev)
an.synth_device(0, W, n, "f32", d_x.data_ptr())
d_rec = torch.zeros((W, 128), dtype=torch.uint8, device=dev)
an.analyze_device(d_x.data_ptr(), W, n, n, "f32", 125.0, d_rec.data_ptr())
torch.cuda.synchronize()
recs = d_rec.cpu().numpy().view(record_dtype(5)).reshape(-1)
odd = np.nonzero(recs["count"] != 3)[0]
print("windows with count != 3:", odd[:20], len(odd))
# compare a sample + the odd ones against fp64 on identical samples
check = np.unique(np.concatenate([odd[:50], np.arange(0, W, max(W // 2000, 1))]))
x = d_x[torch.as_tensor(check, device=dev)].cpu().numpy()
r64 = an.analyze(x.astype(np.float64), 125.0, flexible=True)
diff = 0
for i, w in enumerate(check):
    a = [int(v) for v in recs[w]["pk"]["idx"][: recs[w]["count"]]]
    b = [int(v) for v in r64[i]["pk"]["idx"][: r64[i]["count"]]]
    if a != b:
        diff += 1
        print("window", w, "fp32", prominence_dicts(recs[w], 125.0, n), "\n   fp64", prominence_dicts(r64[i], 125.0, n))
print("checked", len(check), "index mismatches vs fp64:", diff)
