#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ (run in the BUILD container).

  warp_cv2.npz        cv2.warpAffine(diag(s,s), INTER_CUBIC, BORDER_CONSTANT) outputs (rtpose.cpp:484); pins
                      oracle.orc_warp_affine_cubic_u8c3.
  area_cv2.npz        cv2.resize(INTER_AREA) outputs (cv2 4.13 here == the OpenCV algorithm the reference
                      calls at rtpose.cpp:516) for small uint8 BGR images; pins oracle.orc_resize_area_u8c3.
  parse_<model>.npz   reference-pinned stage outputs on a seeded injected scene: stride-8 maps (float16-free,
                      small), peaks blob, joints, subset, JSON text - produced by the oracle AND cross-checked
                      here against the reference's own connectLimbs*/COCO code (oracle/_ref/libref_host.so).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_rtpose_b200 import synth  # noqa: E402
from oracle import orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def gen_area():
    import cv2
    d = {}
    cases = [(90, 160, 46, 82), (90, 160, 40, 70), (72, 128, 48, 96), (64, 96, 32, 48), (96, 96, 32, 48), (45, 80, 23, 41)]
    for i, (sh, sw, dh, dw) in enumerate(cases):
        img = synth.make_frame(100 + i, sh, sw)
        d["src%d" % i] = img
        d["dst%d" % i] = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA)
    d["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(OUT, "area_cv2.npz"), **d)
    print("area_cv2.npz", len(cases), "cases, cv2", cv2.__version__)


def gen_warp():
    import cv2
    d = {}
    cases = [(54, 96, 64, 36), (40, 60, 64, 36), (36, 64, 64, 36), (30, 40, 64, 36), (45, 45, 80, 48)]
    for i, (sh, sw, dw, dh) in enumerate(cases):
        img = synth.make_frame(200 + i, sh, sw)
        s = dw / float(sw) if sw / float(sh) > dw / float(dh) else dh / float(sh)
        M = np.eye(2, 3)
        M[0, 0] = M[1, 1] = s
        d["src%d" % i] = img
        d["dst%d" % i] = cv2.warpAffine(img, M, (dw, dh), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT, borderValue=(0, 0, 0))
        d["scale%d" % i] = np.float64(s)
    np.savez_compressed(os.path.join(OUT, "warp_cv2.npz"), **d)
    print("warp_cv2.npz", len(cases), "cases")


def gen_parse(model, name, net_w, net_h, disp_w, disp_h, n_people, num_scales, seed):
    people = synth.make_people(model, n_people, net_w, net_h, seed=seed)
    maps = synth.make_maps(model, people, net_w, net_h, num_scales=num_scales, start_scale=1.0, scale_gap=0.15, seed=seed)
    thr, p = orc.default_params(model)
    P, mp = orc.num_parts(model), orc.max_peaks(model)
    full = orc.imresize(maps, net_h, net_w, 1.0, 0.15)
    peaks = orc.nms(full, P, mp, thr)
    assert peaks[:, 0, 0].max() <= mp
    cnt, joints, subset = orc.connect(model, full, peaks, disp_w, disp_h, want_subset=True)
    # cross-check with the reference's own code
    p0 = orc.ConnectParams(p.min_subset_cnt, p.min_subset_score, p.inter_threshold, p.inter_min_above, 0)
    c2, j2, s2 = orc.ref_connect(model, full, peaks, disp_w, disp_h, p0)
    assert c2 == cnt and np.array_equal(j2, joints) and np.array_equal(s2, subset), "oracle != reference code"
    js = orc.json_text(joints, P, 1.0)
    np.savez_compressed(os.path.join(OUT, "parse_%s.npz" % name), maps=maps, peaks=peaks, joints=joints, subset=subset,
                        json=np.array(js), meta=np.array([model, net_w, net_h, disp_w, disp_h, num_scales, seed]),
                        start_scale=np.float64(1.0), scale_gap=np.float64(0.15), nms_threshold=np.float32(thr))
    print("parse_%s.npz" % name, "people", cnt, "subset rows", len(subset), "peaks/part max", int(peaks[:, 0, 0].max()))


RENDER_CASES = {"coco": [(0, 1), (1, 0), (19, 0), (20, 0), (25, 0)], "mpi": [(0, 0), (3, 0), (20, 0)]}   # (part_to_show, googly)


def gen_render(name):
    """Renderer regression vectors of the ORACLE (the reference renders on the GPU only; the bit-level pin is
    oracle/_ref/libref_render.so in tests/test_gpu_render.py): flat grey canvas + the persons of parse_<name>.npz."""
    g = np.load(os.path.join(OUT, "parse_%s.npz" % name))
    model, net_w, net_h, disp_w, disp_h, S, _ = [int(v) for v in g["meta"]]
    full = orc.imresize(g["maps"], net_h, net_w, float(g["start_scale"]), float(g["scale_gap"]))
    canvas = np.full((3, disp_h, disp_w), 96.0, np.float32)
    d = {}
    for part, googly in RENDER_CASES[name]:
        out = orc.render(model, canvas, net_w, net_h, full, g["joints"], len(g["joints"]), part, bool(googly))
        d["p%d_g%d" % (part, googly)] = orc.canvas_to_u8(out)
    np.savez_compressed(os.path.join(OUT, "render_%s.npz" % name), **d)
    print("render_%s.npz" % name, {k: int((v != 96).any(2).sum()) for k, v in d.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    gen_area()
    gen_warp()
    gen_parse(orc.COCO_18, "coco", 320, 176, 640, 352, 6, 1, 11)
    gen_parse(orc.COCO_18, "coco_s3", 320, 176, 640, 352, 5, 3, 12)
    gen_parse(orc.MPI_15, "mpi", 240, 176, 480, 352, 4, 1, 13)
    gen_render("coco")
    gen_render("mpi")


if __name__ == "__main__":
    main()
