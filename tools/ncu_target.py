#!/usr/bin/env python3
"""Small, deterministic target for ncu (run under gpurun): one engine, C2 geometry, B frames, N eager forwards (no CUDA graph, one
lane, so that launch k of forward f is launch f*L + k).  tools/ncu_round.sh wraps it."""
import os
import sys

os.environ.setdefault("PE_GRAPH", "0")
os.environ.setdefault("PE_TWO_LANES", "0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_rtpose_b200 import engine, synth  # noqa: E402

B = int(os.environ.get("NT_BATCH", "9"))
N = int(os.environ.get("NT_FORWARDS", "3"))
eng = engine.PoseEngine(engine.COCO_18, 656, 368, 1280, 720, max_batch=B, precision=engine.PREC_F16X2)
eng.set_weights(synth.make_weights(engine.COCO_18, "he"))
frames = [synth.make_frame(i) for i in range(B)]
for _ in range(N):
    eng.forward_frames(frames)
    eng.sync()
print("launches per forward:", eng.launch_count() // N if N else 0)
eng.close()
