#!/usr/bin/env python3
"""Per-layer-class device time of one bench step (run under gpurun): events between the launches of
pe_profile_layers, median of 5 passes, summed per class.  Used for A/B of conv-kernel knobs (PE_TC_* env)."""
import json
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_rtpose_b200 import engine, synth  # noqa: E402


def klass(name, fl):
    if fl == 0:
        return "pool/copy"
    if name in ("conv1_1", "conv1_2"):
        return "conv1_x"
    if name.startswith("conv5_") and name[6] in "123":
        return "stage1 3x3"
    if name.startswith("conv") and not name.startswith("conv5_"):
        return "vgg 3x3"
    if name.startswith("Mconv1_"):
        return "7x7 first"
    if name.startswith("Mconv") and name[5] in "2345":
        return "7x7"
    return "1x1"


def main():
    B = int(os.environ.get("LT_BATCH", "9"))
    model = engine.COCO_18
    eng = engine.PoseEngine(model, 656, 368, 1280, 720, max_batch=B, precision=engine.PREC_BF16X2)
    eng.set_weights(synth.make_weights(model, "he"))
    frames = [synth.make_frame(i) for i in range(B)]
    for _ in range(2):
        eng.forward_frames(frames)
        eng.sync()
    runs = [eng.profile_layers(B) for _ in range(5)]
    names = [n for n, _, _ in runs[0]]
    fl = [f for _, _, f in runs[0]]
    med = [statistics.median(r[i][1] for r in runs) for i in range(len(names))]
    agg, flops = {}, {}
    for n, ms, f in zip(names, med, fl):
        k = klass(n, f)
        agg[k] = agg.get(k, 0.0) + ms
        flops[k] = flops.get(k, 0.0) + f
    tot = sum(agg.values())
    out = {"total_ms": round(tot, 3), "conv_ms": round(tot - agg.get("pool/copy", 0), 3),
           "classes": {k: {"ms": round(v, 3), "tflops": round(flops[k] / v / 1e9, 1) if flops[k] else None} for k, v in sorted(agg.items())}}
    if os.environ.get("LT_VERBOSE"):
        out["layers"] = [(n, round(m, 4)) for n, m in zip(names, med)]
    print(json.dumps(out), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
