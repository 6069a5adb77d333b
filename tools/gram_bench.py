"""Time the tcgen05 Gram kernel (and the SIMT cross-check) on a few shapes; used for the ncu capture in profiles/."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "when-do-gnns-help_b200"))
import wdgh_b200 as W  # noqa: E402

out = []
for m, d in ((10000, 10), (10000, 128), (8192, 1024), (16384, 2048), (500, 1433)):
    z = torch.randn(m, d, device="cuda")
    for tc in (True, False):
        if not tc and m * m * d > 3e11:
            continue
        for _ in range(3):
            W.graph.gram(z, use_tensor_cores=tc)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            g = W.graph.gram(z, use_tensor_cores=tc)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        ref = z.double() @ z.double().T
        err = ((g.double() - ref).abs().max() / ref.abs().max()).item()
        out.append({"m": m, "d": d, "tensor_cores": tc, "ms": round(ms, 4), "useful_tflops": round(2 * m * m * d / ms / 1e9, 1),
                    "executed_tf32_tflops": round(3 * 2 * m * m * d / 2 / ms / 1e9, 1) if tc else None, "rel_err": err})
        print(json.dumps(out[-1]))
