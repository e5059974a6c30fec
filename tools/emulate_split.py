"""CPU emulation of the split-precision conv stack with EXACT (fp64) accumulation: isolates the representational error
of a 2-plane split (bf16 or fp16 planes; products hi*hi + hi*lo + lo*hi) from the tensor core's fp32 accumulate
behaviour.  Usage: python tools/emulate_split.py [net_w net_h]"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as Fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_rtpose_b200 import synth, engine  # noqa: E402


def split(x32, dt):
    hi = x32.to(dt)
    lo = (x32 - hi.to(torch.float32)).to(dt)
    return hi.to(torch.float64), lo.to(torch.float64)


def run(spec, W, x, mode):
    blobs = {"image": x}
    for l in spec["layers"]:
        t = l["type"]
        if t == "Convolution":
            a = blobs[l["bottom"][0]]
            w, b = W[l["name"]]
            w = torch.from_numpy(w)
            b = torch.from_numpy(b).double()
            p = l["pad"]
            if mode == "fp64":
                y = Fn.conv2d(a.double(), w.double(), b, padding=p)
            else:
                dt = torch.bfloat16 if mode == "bf16x2" else torch.float16
                ah, al = split(a, dt)
                wh, wl = split(w, dt)
                y = Fn.conv2d(ah, wh, None, padding=p) + Fn.conv2d(ah, wl, None, padding=p) + Fn.conv2d(al, wh, None, padding=p)
                y = y + b.view(1, -1, 1, 1)
            blobs[l["top"][0]] = y.float() if mode != "fp64" else y
        elif t == "ReLU":
            blobs[l["top"][0]] = torch.relu(blobs[l["bottom"][0]])
        elif t == "Pooling":
            blobs[l["top"][0]] = Fn.max_pool2d(blobs[l["bottom"][0]], 2, 2, ceil_mode=True)
        elif t == "Concat":
            blobs[l["top"][0]] = torch.cat([blobs[n] for n in l["bottom"]], 1)
        elif t in ("ImResize", "Nms"):
            break
    return blobs["concat_stage7"].double()


def main():
    net_w, net_h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (160, 96)
    spec = json.load(open(os.path.join(ROOT, "tests/golden/netspec_coco.json")))
    W = synth.make_weights(engine.COCO_18, "he")
    rng = np.random.default_rng(3)
    x = torch.from_numpy((rng.random((1, 3, net_h, net_w), dtype=np.float32) - 0.5))
    torch.set_num_threads(16)
    ref = run(spec, W, x.double(), "fp64")
    for mode in ("bf16x2", "fp16x2"):
        y = run(spec, W, x, mode)
        print("%s  rel max err %.3e   rel rms err %.3e" % (mode, float((y - ref).abs().max() / ref.abs().max()),
                                                          float(((y - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt())))


if __name__ == "__main__":
    main()
