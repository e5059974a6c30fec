#!/usr/bin/env python3
"""Time the BASELINE.json configs on one GPU (run under gpurun).  Not the bench contract (bench.py = C2); this fills
the results table of DESIGN.md section 5 for C3 (3 scales) and C5's per-GPU workload (992x736, 4 scales)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_rtpose_b200 import engine, synth  # noqa: E402

CONFIGS = [
    ("C1", engine.MPI_15, 496, 368, 640, 480, 1, 1.0, 0.3, 12),
    ("C2", engine.COCO_18, 656, 368, 1280, 720, 1, 1.0, 0.3, 9),
    ("C2-batch1", engine.COCO_18, 656, 368, 1280, 720, 1, 1.0, 0.3, 1),
    ("C3", engine.COCO_18, 656, 368, 1280, 720, 3, 1.0, 0.15, 3),
    ("C5-per-GPU", engine.COCO_18, 992, 736, 1920, 1080, 4, 1.0, 0.15, 1),
]


def main():
    import torch
    out = []
    for name, model, nw, nh, dw, dh, S, start, gap, B in CONFIGS:
        W = synth.make_weights(model, "he")
        engs = [engine.PoseEngine(model, nw, nh, dw, dh, num_scales=S, start_scale=start, scale_gap=gap, max_batch=B,
                                  precision=engine.PREC_BF16X2) for _ in range(2)]
        for e in engs:
            e.set_weights(W)
        frames = torch.from_numpy(np.stack([synth.make_frame(i, dh, dw) for i in range(2 * B)])).cuda()
        fb = dh * dw * 3
        for i in range(4):
            engs[i % 2].forward_frames_device(frames.data_ptr() + (i % 2) * B * fb, B)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 10
        ev0.record()
        for i in range(steps):
            engs[i % 2].forward_frames_device(frames.data_ptr() + (i % 2) * B * fb, B)
        for e in engs:
            e.sync()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        flops = engs[0].conv_flops_per_scale() * S * B
        r = {"config": name, "net": "%dx%d" % (nw, nh), "scales": S, "frames_per_forward": B, "ms_per_forward": ms,
             "frames_per_s": 1e3 * B / ms, "conv_tflops_algorithmic": flops / ms / 1e9}
        print(json.dumps(r), flush=True)
        out.append(r)
        for e in engs:
            e.close()
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
