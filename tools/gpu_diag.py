#!/usr/bin/env python3
"""GPU bring-up diagnostics (run under gpurun): layer-wise error of every conv precision mode against
the oracle at a small net size, post-processing parity, and rough timings.  Prints, never asserts."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_rtpose_b200 import engine, synth  # noqa: E402
from oracle import orc  # noqa: E402

BLOBS = ["image", "conv1_1", "conv1_2", "pool1_stage1", "conv2_2", "conv3_4", "conv4_2",
         "conv4_4_CPM", "conv5_3_CPM_L1", "conv5_4_CPM_L2", "Mconv1_stage2_L1", "Mconv5_stage2_L2", "Mconv6_stage6_L1"]


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def conv_diag(precs, net_w=160, net_h=96, model=engine.COCO_18):
    W = synth.make_weights(model, "he")
    frame = synth.make_frame(0, 2 * net_h, 2 * net_w)
    onet = orc.Net(model)
    onet.set_weights(W)
    x = orc.preprocess(frame, net_h, net_w, 1, 1.0, 0.3)
    ref = {}
    shapes = {}
    t = time.time()
    omaps = onet.forward(x)
    print("oracle forward %.2fs" % (time.time() - t), flush=True)
    for prec in precs:
        try:
            t = time.time()
            eng = engine.PoseEngine(model, net_w, net_h, 2 * net_w, 2 * net_h, precision=prec)
            eng.set_weights(W)
            print("prec %d: create+weights %.2fs" % (prec, time.time() - t), flush=True)
            eng.forward_frames([frame])
            eng.sync()
            maps = eng.fetch_maps(1)
            print("prec %d: stride-8 maps rel err %.3e (abs max %.3e, ref max %.3f)" % (
                prec, rel(maps, omaps), np.abs(maps - omaps).max(), np.abs(omaps).max()), flush=True)
            for b in BLOBS:
                got = eng.fetch_blob(b)[:1]
                if b not in ref:
                    if b == "image":
                        ref[b] = x
                    else:
                        ref[b] = onet.forward_blob(x, b, got.shape[1:])
                print("   %-20s rel err %.3e" % (b, rel(got, ref[b])), flush=True)
            eng.close()
        except Exception:
            traceback.print_exc()


def post_diag():
    for model, net_w, net_h, S in [(engine.COCO_18, 328, 184, 1), (engine.MPI_15, 248, 184, 1), (engine.COCO_18, 328, 184, 3), (engine.COCO_18, 656, 368, 1)]:
        try:
            people = synth.make_people(model, 8, net_w, net_h, seed=2)
            maps8 = synth.make_maps(model, people, net_w, net_h, num_scales=S, start_scale=1.0, scale_gap=0.15, seed=2)
            eng = engine.PoseEngine(model, net_w, net_h, 2 * net_w, 2 * net_h, num_scales=S, start_scale=1.0, scale_gap=0.15,
                                    precision=engine.PREC_FP32_SIMT)
            t = time.time()
            eng.forward_maps(maps8)
            cnt, joints, peaks = eng.fetch(0)
            dt = time.time() - t
            full = orc.imresize(maps8, net_h, net_w, 1.0, 0.15)
            thr, _ = orc.default_params(model)
            opk = orc.nms(full, eng.num_parts, eng.max_peaks, thr)
            ocnt, oj = orc.connect(model, full, opk, 2 * net_w, 2 * net_h)
            print("post model=%d %dx%d S=%d: people %d/%d peaks_equal=%s joints_equal=%s (%.1f ms first call)" % (
                model, net_w, net_h, S, cnt, ocnt, np.array_equal(peaks, opk), cnt == ocnt and np.array_equal(joints, oj), dt * 1e3), flush=True)
            if not np.array_equal(peaks, opk):
                d = np.abs(peaks - opk)
                print("   peaks max diff", d.max(), "counts", peaks[:, 0, 0], opk[:, 0, 0])
            R = orc.ref_cpm()
            if R is not None:
                rfull = np.zeros_like(full)
                R.ref_imresize_host(np.ascontiguousarray(maps8), rfull, S, maps8.shape[1], net_h // 8, net_w // 8, net_h, net_w, 1.0, 0.15)
                rpk = np.zeros_like(opk)
                R.ref_nms_host(rfull, rpk, rfull.shape[0], net_h, net_w, eng.num_parts, eng.max_peaks, thr)
                print("   reference CUDA kernels: resize==oracle %s (max diff %.3e), nms==oracle %s" % (
                    np.array_equal(rfull, full), np.abs(rfull - full).max(), np.array_equal(rpk, opk)), flush=True)
            eng.close()
        except Exception:
            traceback.print_exc()


def timing(prec, batch=1, net_w=656, net_h=368, iters=5):
    try:
        model = engine.COCO_18
        W = synth.make_weights(model, "he")
        eng = engine.PoseEngine(model, net_w, net_h, 1280, 720, precision=prec, max_batch=batch)
        eng.set_weights(W)
        frames = [synth.make_frame(i) for i in range(batch)]
        eng.forward_frames(frames)
        eng.sync()
        eng.event_record(0)
        for _ in range(iters):
            eng.forward_frames(frames)
        eng.event_record(1)
        ms = eng.event_elapsed_ms(0, 1) / iters
        fl = eng.conv_flops_per_scale() * batch
        print("timing prec=%d batch=%d: %.3f ms/forward  %.1f fps  conv %.1f TFLOP/s" % (prec, batch, ms, batch * 1e3 / ms, fl / ms / 1e9), flush=True)
        prof = eng.profile_layers(batch)
        tot = sum(p[1] for p in prof)
        print("   per-layer total %.3f ms; top:" % tot)
        for name, m, f in sorted(prof, key=lambda p: -p[1])[:12]:
            print("     %-22s %.3f ms  %.1f TF/s" % (name, m, f / max(m, 1e-9) / 1e9))
        eng.close()
    except Exception:
        traceback.print_exc()


def render_diag():
    """wall time of pe_render (canvas fill + overlay kernel(s) + uint8 conversion + 2.8 MB D2H) at 1280x720."""
    try:
        model, net_w, net_h = engine.COCO_18, 656, 368
        people = synth.make_people(model, 20, net_w, net_h, seed=4)
        maps8 = synth.make_maps(model, people, net_w, net_h, seed=4)
        eng = engine.PoseEngine(model, net_w, net_h, 1280, 720, precision=engine.PREC_FP32_SIMT)
        eng.forward_maps(maps8)
        cnt, _, _ = eng.fetch(0)
        frame = synth.make_frame(0)
        for part in (0, 5, 19, 20, 25):
            eng.render(0, part, display_bgr=frame)
            t = time.time()
            for _ in range(20):
                eng.render(0, part, display_bgr=frame)
            print("render 1280x720 people=%d part_to_show=%d: %.3f ms per frame (incl. 2.8 MB H2D + 2.8 MB D2H, synchronous)" % (
                cnt, part, (time.time() - t) / 20 * 1e3), flush=True)
        eng.close()
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("post", "all"):
        post_diag()
    if what in ("render",):
        render_diag()
    if what in ("simt", "all"):
        conv_diag([engine.PREC_FP32_SIMT])
    precs = [int(v) for v in os.environ.get("DIAG_PRECS", "1,2").split(",")]
    if what in ("tc", "all"):
        nw, nh = [int(v) for v in os.environ.get("DIAG_NET", "160x96").split("x")]
        conv_diag(precs, nw, nh)
    if what in ("time", "all"):
        for p in precs:
            timing(p, batch=int(os.environ.get("DIAG_BATCH", "8")))
