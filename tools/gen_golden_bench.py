#!/usr/bin/env python3
"""Golden outputs of the ORACLE for the BASELINE.json configurations at FULL size and DEFAULT thresholds
(tests/golden/bench_<cfg>.npz), so that the -m gpu parity tests compare the benched configuration itself
(batch 9, CUDA-graph replay) without spending minutes of CPU time on the GPU box:

  c2   COCO 656x368, 1 scale, frames synth.make_frame(0..8) at 1280x720   (the bench.py step: 9 frames)
  c1   MPI 496x368, 1 scale, synth.make_frame(42, 480, 640)
  c3   COCO 656x368, 3 scales (1.0 / 0.85 / 0.70), synth.make_frame(42)
  c5   COCO 992x736, 4 scales (gap 0.15), synth.make_frame(5, 1080, 1920)

Stored per frame: person count, joints, peaks blob, and the stride-8 maps (all channels for frame 0 of c2, a
channel subset elsewhere - the maps are noise and do not compress).  W-he weights, seed 1234 (synth.make_weights).
Run in the build container:  python tools/gen_golden_bench.py [c2 c1 c3 c5]     (~8 minutes on 8 cores)
tests/test_oracle.py re-runs one c2 frame live against the file (the golden IS the oracle on this machine)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_rtpose_b200 import synth  # noqa: E402
from oracle import orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
MAP_SUBSET = [0, 7, 18, 19, 38, 56]      # parts, background, PAFs (COCO channel numbering; MPI uses [0, 7, 15, 16, 30, 43])
MAP_SUBSET_MPI = [0, 7, 15, 16, 30, 43]

CONFIGS = {
    "c2": dict(model=orc.COCO_18, net=(656, 368), disp=(1280, 720), S=1, start=1.0, gap=0.3, frames=[(i, 720, 1280) for i in range(9)]),
    "c1": dict(model=orc.MPI_15, net=(496, 368), disp=(640, 480), S=1, start=1.0, gap=0.3, frames=[(42, 480, 640)]),
    "c3": dict(model=orc.COCO_18, net=(656, 368), disp=(1280, 720), S=3, start=1.0, gap=0.15, frames=[(42, 720, 1280)]),
    "c5": dict(model=orc.COCO_18, net=(992, 736), disp=(1920, 1080), S=4, start=1.0, gap=0.15, frames=[(5, 1080, 1920)]),
}


def gen(name):
    c = CONFIGS[name]
    model = c["model"]
    net = orc.Net(model)
    net.set_weights(synth.make_weights(model, "he"))
    net_w, net_h = c["net"]
    sub = MAP_SUBSET if model == orc.COCO_18 else MAP_SUBSET_MPI
    d = {"frames": np.array(c["frames"]), "net": np.array(c["net"]), "disp": np.array(c["disp"]), "S": np.int32(c["S"]),
         "start": np.float64(c["start"]), "gap": np.float64(c["gap"]), "model": np.int32(model), "map_subset": np.array(sub)}
    for i, (seed, h, w) in enumerate(c["frames"]):
        t = time.time()
        frame = synth.make_frame(seed, h, w)
        cnt, joints, peaks, maps = net.process_frame(frame, net_h, net_w, c["S"], c["start"], c["gap"])
        d["cnt%d" % i] = np.int32(cnt)
        d["joints%d" % i] = joints
        d["peaks%d" % i] = peaks
        d["maps_sub%d" % i] = maps[:, sub]
        d["maps_absmax%d" % i] = np.float32(np.abs(maps).max())
        if name == "c2" and i == 0:
            d["maps_full0"] = maps
        print("%s frame %d: %d people, %d peaks (%.1f s)" % (name, i, cnt, int(np.minimum(peaks[:, 0, 0], peaks.shape[1] - 1).sum()),
                                                            time.time() - t), flush=True)
    np.savez_compressed(os.path.join(OUT, "bench_%s.npz" % name), **d)


if __name__ == "__main__":
    for n in (sys.argv[1:] or ["c2", "c1", "c3", "c5"]):
        gen(n)
