"""profiles/r02_bf16_parity.md from the dump of tests/test_bf16_parity_gpu.py (COMBAT_PARITY_DUMP=gpurun_out/parity.jsonl)."""
import json
import sys

import numpy as np

rows = [json.loads(l) for l in open(sys.argv[1])]
print("# r02: per-tensor parity of the bf16 / tcgen05 path (B200, tests/test_bf16_parity_gpu.py)\n")
print("Columns: `CUDA vs Q` = relative L2 distance of the CUDA bf16 path from the quantisation-aware oracle (`with O.quantised():`);")
print("`floor` = distance of the quantised oracle from ITSELF when every conv output carries 2e-6 relative noise")
print("(`quantised(jitter=2e-6)`, worst of two seeds) -- what float32 accumulation-order noise alone does to this algorithm;")
print("`CUDA vs fp32` / `Q vs fp32` = distances from the float32 reference restatement.  Every CUDA conv kernel sits 5e-7 .. 4e-6")
print("from torch's float32 conv on identical inputs (`test_*_layers_match_torch_on_their_own_input`), so `CUDA vs Q <= ~floor`")
print("means the implementation is indistinguishable from a re-run of its own algorithm.\n")
for r in rows:
    print("## %s\n" % r["case"])
    print("| tensor | CUDA vs Q | floor | CUDA vs fp32 | Q vs fp32 |\n|---|---:|---:|---:|---:|")
    for x in r["rows"]:
        if x[0] in ("fwd", "loss"):
            print("| %s | %.2e | %.2e | %.2e | %.2e |" % (x[1], x[2], x[3], x[4], x[5]))
    g = [x for x in r["rows"] if x[0] == "grad"]
    e, f = np.array([x[2] for x in g]), np.array([x[3] for x in g])
    c, cf, b = np.array([x[4] for x in g]), np.array([x[5] for x in g]), np.array([x[6] for x in g])
    print("\nParameter gradients, %d tensors (every weight / affine tensor of netC and netG; InstanceNorm-dead biases excluded):\n" % len(g))
    print("| | median | max | tensor at the max |\n|---|---:|---:|---|")
    print("| CUDA vs Q (L2 rel) | %.3f | %.3f | %s |" % (np.median(e), e.max(), g[int(e.argmax())][1]))
    print("| floor (Q vs jittered Q) | %.3f | %.3f | %s |" % (np.median(f), f.max(), g[int(f.argmax())][1]))
    print("| CUDA vs Q / floor | %.2f | %.2f | %s |" % (np.median(e / f), (e / f).max(), g[int((e / f).argmax())][1]))
    print("| Q vs fp32 (cost of bf16 storage) | %.3f | %.3f | %s |" % (np.median(b), b.max(), g[int(b.argmax())][1]))
    print("| cosine CUDA vs Q (min) | | %.4f | %s |" % (c.min(), g[int(c.argmin())][1]))
    print("| cosine floor (min) | | %.4f | %s |\n" % (cf.min(), g[int(cf.argmin())][1]))
