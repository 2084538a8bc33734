"""GPU bring-up probe for the tcgen05 GEMM: exact-integer problems whose result pins operand layout mistakes
(swizzle, descriptor strides, TMEM lane/column mapping).  Writes gpurun_out/gemm_diag.{json,npz}."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from affectgpt_b200 import ops  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)
dev = "cuda:0"
res, dump = [], {}
g = torch.Generator().manual_seed(0)
for (M, N, K) in [(128, 256, 64), (128, 256, 128), (128, 256, 1024), (256, 512, 64), (100, 64, 64), (300, 768, 256), (2048, 256, 64), (2056, 512, 128), (4096, 768, 1024)]:
    a = torch.randint(-2, 3, (M, K), generator=g).to(torch.bfloat16).to(dev)
    w = torch.randint(-2, 3, (N, K), generator=g).to(torch.bfloat16).to(dev)
    try:
        of, ob = ops.gemm(a, w, None, out_f32=True, out_bf16=True)
        torch.cuda.synchronize()
        want = a.float() @ w.float().T
        bad = (of != want)
        entry = {"shape": [M, N, K], "exact": bool(not bad.any()), "mismatches": int(bad.sum()),
                 "bad_rows": int(bad.any(1).sum()), "bad_cols": int(bad.any(0).sum()),
                 "first_bad": [int(x) for x in bad.nonzero()[0]] if bad.any() else None,
                 "bf16_exact": bool((ob.float() == want).all())}
        if bad.any() and len(dump) < 4:
            dump[f"got_{M}_{N}_{K}"] = of.cpu().numpy()
            dump[f"want_{M}_{N}_{K}"] = want.cpu().numpy()
            dump[f"a_{M}_{N}_{K}"] = a.float().cpu().numpy()
            dump[f"w_{M}_{N}_{K}"] = w.float().cpu().numpy()
    except Exception as e:  # noqa: BLE001
        entry = {"shape": [M, N, K], "error": repr(e)[:400]}
    print(entry, flush=True)
    res.append(entry)
json.dump(res, open(os.path.join(OUT, "gemm_diag.json"), "w"), indent=1)
if dump:
    np.savez_compressed(os.path.join(OUT, "gemm_diag.npz"), **dump)
sys.exit(0 if all(r.get("exact") for r in res) else 1)
