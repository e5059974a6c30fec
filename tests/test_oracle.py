"""CPU tests that PIN THE ORACLE before anything trusts it:

  * graph table            == tests/golden/netspec_*.json parsed from the reference prototxts
  * model descriptors      == the reference's own modelDescriptorFactory.cpp (oracle/_ref, when built)
  * im2col                 == the reference's own im2col_cpu (oracle/_ref)                    bit-exact
  * connectLimbs / COCO    == the reference's own functions (oracle/_ref) on seeded scenes    bit-exact
  * MAX pooling            == upstream Caffe known-answer vector (test_pooling_layer.cpp:49-120)
  * convolution            vs a naive direct loop at 1e-4 (the bar of test_convolution_layer.cpp:231-265)
  * INTER_AREA             == committed cv2 fixtures (tests/golden/area_cv2.npz)               bit-exact
  * stage-level goldens    == tests/golden/parse_*.npz (peaks, joints, subset, JSON)          bit-exact
ImResize/NMS are pinned against the reference's own CUDA kernels in tests/test_gpu_post.py (test_reference_cuda_kernels_equal_oracle).
"""
import json
import os

import numpy as np
import pytest

from caffe_rtpose_b200 import synth
from oracle import orc

MODELS = [(orc.COCO_18, "coco"), (orc.MPI_15, "mpi")]


@pytest.mark.parametrize("model,name", MODELS)
def test_graph_matches_prototxt_fixture(model, name, golden_dir):
    spec = json.load(open(os.path.join(golden_dir, "netspec_%s.json" % name)))
    mine = orc.Net(model).layers()
    assert len(mine) == len(spec["layers"]) == 183
    for a, b in zip(mine, spec["layers"]):
        assert a["name"] == b["name"] and a["type"] == b["type"]
        assert a["bottom"] == b["bottom"] and [a["top"]] == b["top"]
        if a["type"] == "Convolution":
            assert (a["num_output"], a["kernel_size"], a["pad"], a["stride"]) == (
                b["num_output"], b["kernel_size"], b["pad"], b["stride"])
            assert b["weight_filler"] == {"type": "gaussian", "std": "0.01"}
        if a["type"] == "Pooling":
            assert (a["kernel_size"], a["stride"], a["pad"], b["pool"]) == (b["kernel_size"], b["stride"], b["pad"], "MAX")
    nms = spec["layers"][-1]
    assert nms["num_parts"] == orc.num_parts(model) and nms["max_peaks"] == orc.max_peaks(model)
    assert spec["layers"][-2]["factor"] == 8.0
    # conv table used by the weight generator == oracle's
    assert orc.Net(model).convs() == synth.conv_table(model)


def test_flops_match_baseline():
    assert orc.flops(orc.COCO_18, 368, 656) == 484634285056.0
    assert orc.flops(orc.COCO_18, 736, 992) == 1465723203584.0
    assert orc.flops(orc.MPI_15, 368, 496) == 361694564352.0


@pytest.mark.parametrize("model,name", MODELS)
def test_model_descriptor_vs_reference_code(model, name):
    R = orc.ref_host()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference)")
    import ctypes as C
    npart, nlimb = C.c_int(), C.c_int()
    ls, mi = np.zeros(64, np.int32), np.zeros(64, np.int32)
    names = C.create_string_buffer(8192)
    assert R.ref_model_descriptor(model, C.byref(npart), C.byref(nlimb), ls, mi, names, 8192) == 0
    assert npart.value == orc.num_parts(model)
    assert list(ls[:2 * nlimb.value]) == orc.limb_seq(model) == synth._LIMBS[model]
    assert list(mi[:2 * nlimb.value]) == orc.map_idx(model) == synth._MAPIDX[model]
    ref_names = names.value.decode().split("\n")[:-1]
    assert ref_names == [orc.lib().orc_model_map_name(model, i).decode() for i in range(orc.num_maps(model))]


def test_im2col_vs_reference_code():
    R = orc.ref_host()
    if R is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(0)
    for (c, h, w, k, pad) in [(3, 7, 9, 3, 1), (5, 11, 6, 7, 3), (4, 5, 5, 1, 0)]:
        im = rng.standard_normal((c, h, w)).astype(np.float32)
        col = orc.im2col(im, k, pad)
        ref = np.empty_like(col)
        R.ref_im2col(im, c, h, w, k, k, pad, pad, 1, 1, ref)
        assert np.array_equal(col, ref)


def test_maxpool_known_answer_upstream():
    # test_pooling_layer.cpp:49-120 (kernel 2, stride 1)
    x = np.tile(np.array([[1, 2, 5, 2, 3], [9, 4, 1, 4, 8], [1, 2, 5, 2, 3]], np.float32), (2, 2, 1, 1))
    y = orc.maxpool(x, 2, 1, 0)
    assert y.shape == (2, 2, 2, 4)
    assert np.array_equal(y, np.tile(np.array([[9, 5, 5, 8], [9, 5, 5, 8]], np.float32), (2, 2, 1, 1)))


def test_maxpool_ceil_dims():
    # pooling_layer.cpp:90-93: ceil -> odd sizes keep the last partial window
    x = np.arange(2 * 5 * 7, dtype=np.float32).reshape(1, 2, 5, 7)
    y = orc.maxpool(x, 2, 2, 0)
    assert y.shape == (1, 2, 3, 4)
    assert y[0, 0, 2, 3] == x[0, 0, 4, 6] and y[0, 0, 0, 0] == x[0, 0, 1, 1]


def _naive_conv(x, w, b, pad):
    n, cin, h, ww = x.shape
    cout, _, k, _ = w.shape
    xp = np.zeros((n, cin, h + 2 * pad, ww + 2 * pad), np.float64)
    xp[:, :, pad:pad + h, pad:pad + ww] = x
    out = np.zeros((n, cout, h + 2 * pad - k + 1, ww + 2 * pad - k + 1), np.float64)
    for o in range(cout):
        for y in range(out.shape[2]):
            for xx in range(out.shape[3]):
                out[:, o, y, xx] = (xp[:, :, y:y + k, xx:xx + k] * w[o]).sum((1, 2, 3)) + b[o]
    return out


@pytest.mark.parametrize("k,pad", [(3, 1), (1, 0), (7, 3)])
def test_conv_vs_naive_loop(k, pad):
    # upstream bar: 1e-4 vs caffe_conv (test_convolution_layer.cpp:231-265, 443-468), Gaussian-filled 2x3x6x4
    rng = np.random.default_rng(k)
    x = rng.standard_normal((2, 3, 6, 4)).astype(np.float32)
    w = rng.standard_normal((4, 3, k, k)).astype(np.float32)
    b = rng.standard_normal(4).astype(np.float32)
    assert np.abs(orc.conv2d(x, w, b, pad) - _naive_conv(x, w, b, pad)).max() < 1e-4


def test_conv_vs_reference_code_same_blas():
    """ConvolutionLayer::Forward_cpu -> forward_cpu_gemm / forward_cpu_bias -> caffe_cpu_gemm (conv_layer.cpp:27-39,
    base_conv_layer.cpp:259-271, 277-279, math_functions.cpp:12-21) compiled from the reference, its cblas_sgemm bound to the SAME
    OpenBLAS the oracle loads: identical calls into an identical library, so the outputs must be bit-identical - including the
    is_1x1_ shortcut, the bias-as-gemm form and batches.  (The BLAS binary itself is third-party and unpinned in the reference.)"""
    R = orc.ref_host()
    blas = orc.find_blas()
    if R is None or not hasattr(R, "ref_conv_forward") or blas is None or not orc.lib().orc_have_blas():
        pytest.skip("needs oracle/_ref (built from /root/reference) and an OpenBLAS")
    assert R.ref_load_blas(blas.encode()) == 0
    rng = np.random.default_rng(9)
    for (n, cin, h, w, cout, k, pad, with_bias) in [(2, 3, 6, 4, 4, 3, 1, True), (1, 64, 23, 41, 64, 3, 1, True), (2, 128, 12, 21, 128, 7, 3, True),
                                                      (1, 128, 12, 21, 512, 1, 0, True), (1, 185, 12, 21, 128, 7, 3, True)]:
        x = rng.standard_normal((n, cin, h, w)).astype(np.float32)
        wt = (rng.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)
        b = rng.standard_normal(cout).astype(np.float32)
        oh, ow = h + 2 * pad - k + 1, w + 2 * pad - k + 1
        ref = np.full((n, cout, oh, ow), 3.0, np.float32)
        assert R.ref_conv_forward(x, n, cin, h, w, wt, b.ctypes.data, cout, k, pad, ref) == 0
        got = orc.conv2d(x, wt, b, pad)
        assert np.array_equal(got, ref), (n, cin, h, w, cout, k, pad, float(np.abs(got - ref).max()))


def test_relu():
    x = np.array([-1.5, 0.0, 2.0, -0.0], np.float32)
    orc.lib().orc_relu(x, x.size)
    assert np.array_equal(x, np.array([0, 0, 2, 0], np.float32))


def _ref_host2():
    R = orc.ref_host()
    if R is None or not hasattr(R, "ref_maxpool"):
        pytest.skip("oracle/_ref not built (no /root/reference)")
    return R


def test_pool_and_relu_vs_reference_code():
    """PoolingLayer::Reshape + Forward_cpu MAX (pooling_layer.cpp:90-105, 151-186) and ReLULayer::Forward_cpu (relu_layer.cpp:15-18)
    compiled from the reference: sizes (ceil mode, clipped last window), first-maximum semantics, -0.0 / NaN-free inputs."""
    R = _ref_host2()
    rng = np.random.default_rng(3)
    for (n, c, h, w, k, s, pad) in [(2, 3, 8, 10, 2, 2, 0), (1, 4, 7, 9, 2, 2, 0), (1, 2, 46, 82, 2, 2, 0), (1, 2, 9, 9, 3, 2, 1), (2, 1, 5, 5, 3, 2, 0)]:
        x = rng.standard_normal((n, c, h, w)).astype(np.float32)
        x[0, 0, :2, :2] = 0.5          # ties inside one window
        got = orc.maxpool(x, k, s, pad)
        hw = np.zeros(2, np.int32)
        R.ref_maxpool(x, n, c, h, w, k, s, pad, None, hw)
        assert got.shape == (n, c, hw[0], hw[1])
        ref = np.empty_like(got)
        R.ref_maxpool(x, n, c, h, w, k, s, pad, ref.ctypes.data, hw)
        assert np.array_equal(got, ref)
    x = rng.standard_normal(4099).astype(np.float32)
    x[:3] = (0.0, -0.0, -1e-38)
    ref = np.empty_like(x)
    R.ref_relu(x, ref, x.size, 0.0)
    y = x.copy()
    orc.lib().orc_relu(y, y.size)
    assert np.array_equal(y, ref)


def test_preprocess_vs_reference_code():
    """process_and_pad_image (rtpose.cpp:239-269), the display scale (:474-479) and the per-scale target size (:509-511) compiled from
    the reference; the INTER_AREA resize between them is OpenCV's (pinned to cv2 by the fixtures above)."""
    R = _ref_host2()
    import ctypes as C
    for (nw, nh, start, gap, S) in [(656, 368, 1.0, 0.3, 3), (656, 368, 1.0, 0.15, 4), (496, 368, 1.0, 0.3, 2), (160, 96, 1.0, 0.3, 3), (992, 736, 1.0, 0.15, 4),
                                    (656, 368, 0.9, 0.05, 6)]:
        for i in range(S):
            tw, th = C.c_int(), C.c_int()
            R.ref_scale_target(nw, nh, start, gap, i, C.byref(tw), C.byref(th))
            assert orc.scale_target(nw, nh, start, gap, i) == (tw.value, th.value)
    for (cols, rows, dw, dh) in [(1280, 720, 1280, 720), (640, 480, 1280, 720), (1920, 1080, 1280, 720), (333, 777, 656, 368), (1000, 10, 64, 64)]:
        assert orc.lib().orc_display_scale(cols, rows, dw, dh) == R.ref_display_scale(cols, rows, dw, dh)
    img = synth.make_frame(2, 90, 160)
    net_h, net_w, S, start, gap = 48, 96, 3, 1.0, 0.3
    out = orc.preprocess(img, net_h, net_w, S, start, gap)
    for i in range(S):
        tw, th = orc.scale_target(net_w, net_h, start, gap, i)
        small = orc.resize_area(img, th, tw)
        ref = np.full((3, net_h, net_w), 7.0, np.float32)
        R.ref_process_and_pad_image(ref, small, tw, th, net_w, net_h, 1)
        assert np.array_equal(out[i], ref)
    # normalize = 0: the float canvas of the renderers (rtpose.cpp:499)
    ref = np.empty((3, 90, 160), np.float32)
    R.ref_process_and_pad_image(ref, img, 160, 90, 160, 90, 0)
    assert np.array_equal(orc.canvas_from_u8(img), ref)


def test_json_writer_vs_reference_code(tmp_path):
    """The JSON block of displayFrame (rtpose.cpp:1395-1414, `fs << double` formatting) compiled from the reference, byte for byte -
    for the oracle AND for the product's pe_write_json."""
    R = _ref_host2()
    from caffe_rtpose_b200 import engine
    rng = np.random.default_rng(5)
    for people, parts, scale in [(0, 18, 1.0), (1, 18, 0.5), (3, 15, 1.0), (7, 18, 0.3333333), (2, 18, 2.25)]:
        j = (rng.random((people, parts, 3)) * np.array([1280, 720, 1])).astype(np.float32)
        if people:
            j[0, 1] = 0.0                                    # a missing part
            j[0, 2] = (1e-5, 123456.7, 1.0)                  # exponent and 6-digit rounding cases of operator<<(double/float)
        path = str(tmp_path / "ref.json")
        R.ref_write_json(path.encode(), j if j.size else np.zeros(1, np.float32), people, parts, scale)
        want = open(path).read()
        assert orc.json_text(j, parts, scale) == want
        assert engine.write_json(j, parts, scale) == want


def test_model_default_thresholds_vs_reference_code():
    """warmup()'s model selection (rtpose.cpp:212-229) compiled from the reference: the NMS / connect thresholds each model starts with."""
    R = orc.ref_host()
    if R is None or not hasattr(R, "ref_model_defaults"):
        pytest.skip("oracle/_ref not built (no /root/reference)")
    import ctypes as C
    R.ref_model_defaults.argtypes = [C.c_int] + [C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
    for model, parts in ((orc.MPI_15, 15), (orc.COCO_18, 18)):
        thr, cnt, score, inter, above = C.c_float(), C.c_int(), C.c_float(), C.c_float(), C.c_int()
        R.ref_model_defaults(parts, C.byref(thr), C.byref(cnt), C.byref(score), C.byref(inter), C.byref(above))
        othr, p = orc.default_params(model)
        assert (othr, p.min_subset_cnt, p.min_subset_score, p.inter_threshold, p.inter_min_above) == (thr.value, cnt.value, score.value, inter.value, above.value)


def test_render_dispatch_vs_reference_code():
    """render() (rtpose.cpp:271-300) compiled from the reference with recording launchers: for every --part_to_show value of both
    models (and past the last view) the oracle picks the same launcher with the same `part`, googly / num_parts_accum arguments."""
    R = _ref_host2()
    if not hasattr(R, "ref_render_dispatch"):
        pytest.skip("oracle/_ref predates the render() splice")
    for model, parts in ((orc.MPI_15, 15), (orc.COCO_18, 18)):
        for p2s in range(0, 48):
            for googly in (0, 1):
                ref, got = np.zeros(3, np.int32), np.zeros(3, np.int32)
                assert R.ref_render_dispatch(parts, p2s, googly, ref) == 1
                orc.lib().orc_render_dispatch(model, p2s, googly, got)
                assert list(got) == list(ref), (model, p2s, googly)


def test_inter_area_vs_cv2_fixture(golden_dir):
    d = np.load(os.path.join(golden_dir, "area_cv2.npz"))
    n = len([k for k in d.files if k.startswith("src")])
    assert n >= 6
    for i in range(n):
        dst = d["dst%d" % i]
        assert np.array_equal(orc.resize_area(d["src%d" % i], dst.shape[0], dst.shape[1]), dst)


def test_inter_area_live_cv2():
    cv2 = pytest.importorskip("cv2")
    img = synth.make_frame(3, 180, 320)
    # (184, 248): one axis enlarges -> OpenCV's fixed-point bilinear "area mode"; (200, 400): both enlarge
    for dh, dw in [(92, 164), (80, 140), (90, 160), (60, 160), (184, 248), (200, 400)]:
        assert np.array_equal(orc.resize_area(img, dh, dw), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA))


def test_scale_targets():
    # 656x368 @ {1, .85, .70} -> 656x368, 560x320, 464x272
    assert [orc.scale_target(656, 368, 1.0, 0.15, i) for i in range(3)] == [(656, 368), (560, 320), (464, 272)]
    assert [synth.scale_geometry(656, 368, 1.0, 0.15, i)[:2] for i in range(3)] == [(656, 368), (560, 320), (464, 272)]


def test_preprocess_pad_and_normalise():
    img = synth.make_frame(1, 90, 160)
    out = orc.preprocess(img, 48, 96, 2, 1.0, 0.3)
    assert out.shape == (2, 3, 48, 96)
    r0 = orc.resize_area(img, 48, 96)
    assert np.array_equal(out[0], (r0.transpose(2, 0, 1).astype(np.float32) / np.float32(256) - np.float32(0.5)))
    tw, th = orc.scale_target(96, 48, 1.0, 0.3, 1)
    assert (tw, th) == (80, 48) or (tw, th) == (80, 34 + 14)  # 16*ceil(67.2/16)=80, 16*ceil(33.6/16)=48
    padw = (96 - tw) // 2
    assert np.all(out[1][:, :, :padw] == 0) and np.all(out[1][:, :, padw + tw:] == 0)


@pytest.mark.parametrize("name", ["coco", "coco_s3", "mpi"])
def test_stage_goldens(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "parse_%s.npz" % name))
    model, net_w, net_h, disp_w, disp_h, S, _ = [int(v) for v in g["meta"]]
    full = orc.imresize(g["maps"], net_h, net_w, float(g["start_scale"]), float(g["scale_gap"]))
    peaks = orc.nms(full, orc.num_parts(model), orc.max_peaks(model), float(g["nms_threshold"]))
    assert np.array_equal(peaks, g["peaks"])
    cnt, joints, subset = orc.connect(model, full, peaks, disp_w, disp_h, want_subset=True)
    assert cnt == len(g["joints"]) and cnt >= 3
    assert np.array_equal(joints, g["joints"]) and np.array_equal(subset, g["subset"])
    assert orc.json_text(joints, orc.num_parts(model)) == str(g["json"])


@pytest.mark.parametrize("model,net_w,net_h,n", [(orc.COCO_18, 320, 176, 8), (orc.MPI_15, 240, 176, 5), (orc.COCO_18, 656, 368, 22)])
def test_connect_vs_reference_code(model, net_w, net_h, n):
    if orc.ref_host() is None:
        pytest.skip("oracle/_ref not built")
    for seed in range(3):
        people = synth.make_people(model, n, net_w, net_h, seed=seed, drop_prob=0.2)
        maps = synth.make_maps(model, people, net_w, net_h, seed=seed)
        full = orc.imresize(maps, net_h, net_w, 1.0, 0.3)
        thr, p = orc.default_params(model)
        peaks = orc.nms(full, orc.num_parts(model), orc.max_peaks(model), thr)
        assert peaks[:, 0, 0].max() <= orc.max_peaks(model)
        cnt, joints, subset = orc.connect(model, full, peaks, 2 * net_w, 2 * net_h, want_subset=True)
        p0 = orc.ConnectParams(p.min_subset_cnt, p.min_subset_score, p.inter_threshold, p.inter_min_above, 0)
        c2, j2, s2 = orc.ref_connect(model, full, peaks, 2 * net_w, 2 * net_h, p0)
        assert cnt == c2 and cnt >= n // 2
        assert np.array_equal(joints, j2) and np.array_equal(subset, s2)


def test_connect_special_cases_vs_reference_code():
    """nA==0 / nB==0 singleton rows, duplicate check (COCO only), nothing at all."""
    if orc.ref_host() is None:
        pytest.skip("oracle/_ref not built")
    for model, net_w, net_h in [(orc.COCO_18, 320, 176), (orc.MPI_15, 240, 176)]:
        P, mp = orc.num_parts(model), orc.max_peaks(model)
        thr, p = orc.default_params(model)
        p0 = orc.ConnectParams(p.min_subset_cnt, p.min_subset_score, p.inter_threshold, p.inter_min_above, 0)
        people = synth.make_people(model, 5, net_w, net_h, seed=5, drop_prob=0.0)
        for drop in ([2, 3, 4], [1], list(range(P)), [0, 14, 15, 16, 17][:3]):
            ppl = [{k: v for k, v in q.items() if k not in drop} for q in people]
            maps = synth.make_maps(model, ppl, net_w, net_h, seed=1)
            full = orc.imresize(maps, net_h, net_w, 1.0, 0.3)
            peaks = orc.nms(full, P, mp, thr)
            a = orc.connect(model, full, peaks, net_w, net_h, want_subset=True)
            b = orc.ref_connect(model, full, peaks, net_w, net_h, p0)
            assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_nms_quirks():
    """strict >, border exclusion, score>0 filter, width-for-height window bound (nms_layer.cu:79)."""
    H, W = 20, 40
    m = np.zeros((2, H, W), np.float32)
    m[0, 5, 5] = 1.0
    m[0, 5, 6] = 1.0            # tie with neighbour -> neither is a peak (strict >)
    m[0, 10, 10] = 0.9
    m[0, 9, 10] = -5.0          # negative neighbour ignored in the centroid (score > 0)
    m[0, 10, 11] = 0.3
    m[0, 0, 20] = 2.0           # border row: never a peak
    m[0, H - 2, 30] = 0.8       # bottom interior row: window rows H..H+1 alias channel 1 rows 0..1
    m[1, 0, 30] = 0.4
    pk = orc.nms(m, 1, 8, 0.05)
    assert pk[0, 0, 0] == 2
    x, y, s = pk[0, 1]
    assert s == np.float32(0.9) and y == 10 and abs(x - (10 * 0.9 + 11 * 0.3) / 1.2) < 1e-6
    x, y, s = pk[0, 2]
    assert s == np.float32(0.8) and abs(y - ((H - 2) * 0.8 + H * 0.4) / 1.2) < 1e-5  # aliased row pulled y down


def test_nms_count_unclamped_and_first_max_peaks_kept():
    H, W = 16, 64
    m = np.zeros((2, H, W), np.float32)
    for i in range(10):
        m[0, 3 + (i % 2) * 6, 4 + 5 * i] = 0.5 + 0.01 * i
    pk = orc.nms(m, 1, 4, 0.05)
    assert pk[0, 0, 0] == 10                     # total, not clamped (nms_layer.cu:110)
    assert [int(round(v)) for v in pk[0, 1:, 1]] == [3, 3, 3, 3]  # raster order: the y=3 row first


def test_json_format():
    j = np.zeros((1, 18, 3), np.float32)
    j[0, 0] = (618.56, 289.597, 0.950805)
    t = orc.json_text(j, 18, 0.5)
    assert t.startswith('{\n"version":0.1,\n"bodies":[\n{\n"joints":[1237.12,579.194,0.950805,0,0,0,')
    assert t.endswith("]\n}]\n}\n")
    assert orc.json_text(np.zeros((0, 18, 3), np.float32), 18) == '{\n"version":0.1,\n"bodies":[\n]\n}\n'


def test_warp_affine_vs_cv2_fixture(golden_dir):
    d = np.load(os.path.join(golden_dir, "warp_cv2.npz"))
    n = len([k for k in d.files if k.startswith("src")])
    assert n >= 5
    for i in range(n):
        dst = d["dst%d" % i]
        out, s = orc.display_image(d["src%d" % i], dst.shape[1], dst.shape[0])
        assert s == float(d["scale%d" % i]) and np.array_equal(out, dst)


def test_warp_affine_live_cv2():
    cv2 = pytest.importorskip("cv2")
    for (sh, sw, dw, dh) in [(108, 192, 128, 72), (48, 64, 128, 72), (72, 128, 128, 72), (100, 100, 128, 72)]:
        img = synth.make_frame(5, sh, sw)
        out, s = orc.display_image(img, dw, dh)
        M = np.eye(2, 3)
        M[0, 0] = M[1, 1] = s
        ref = cv2.warpAffine(img, M, (dw, dh), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT, borderValue=(0, 0, 0))
        assert np.array_equal(out, ref)
    same, s = orc.display_image(synth.make_frame(1, 72, 128), 128, 72)   # scale 1: identity
    assert s == 1.0 and np.array_equal(same, synth.make_frame(1, 72, 128))


# ---- renderers (render() rtpose.cpp:271-300, renderFunctions.cu): GPU-only in the reference, so the CPU suite holds
# regression vectors of the restatement + its invariants; the bit-level pin against the reference's own kernels
# (oracle/_ref/libref_render.so) is tests/test_gpu_render.py.
@pytest.mark.parametrize("name", ["coco", "mpi"])
def test_render_goldens(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "parse_%s.npz" % name))
    r = np.load(os.path.join(golden_dir, "render_%s.npz" % name))
    model, net_w, net_h, disp_w, disp_h, S, _ = [int(v) for v in g["meta"]]
    full = orc.imresize(g["maps"], net_h, net_w, float(g["start_scale"]), float(g["scale_gap"]))
    canvas = np.full((3, disp_h, disp_w), 96.0, np.float32)
    for key in r.files:
        part, googly = int(key.split("_")[0][1:]), int(key.split("_")[1][1:])
        img = orc.canvas_to_u8(orc.render(model, canvas, net_w, net_h, full, g["joints"], len(g["joints"]), part, bool(googly)))
        assert (img != r[key]).any(2).mean() < 1e-4, key   # libm sinf/cosf may move a border pixel between glibc builds


def test_render_invariants():
    model, w, h = orc.COCO_18, 96, 64
    canvas = orc.canvas_from_u8(synth.make_frame(3, h, w))
    joints = np.zeros((1, 18, 3), np.float32)
    # no people / no confident joint: the skeleton view leaves the canvas untouched (renderFunctions.cu:1006, :437)
    assert np.array_equal(orc.render(model, canvas, 48, 32, None, joints, 0, 0), canvas)
    assert np.array_equal(orc.render(model, canvas, 48, 32, None, joints, 1, 0), canvas)
    # one limb (neck-right shoulder): an ellipse around the segment in the limb's colour, alpha 0.5, plus two joint discs
    joints[0, 1] = (30, 30, 1.0)
    joints[0, 2] = (60, 30, 1.0)
    out = orc.render(model, canvas, 48, 32, None, joints, 1, 0)
    changed = np.argwhere((out != canvas).any(0))
    assert len(changed) > 0 and changed[:, 1].min() >= 29 and changed[:, 1].max() <= 61 and abs(changed[:, 0].mean() - 30) < 1
    mid = out[:, 30, 45]
    assert np.allclose(mid, 0.5 * canvas[:, 30, 45] + 0.5 * np.array([0, 0, 255], np.float32))   # colour 0 = (r 255, g 0, b 0)
    # float canvas -> uint8: int(v + 0.5) with clamping (rtpose.cpp:1291-1293)
    c = np.zeros((3, 1, 4), np.float32)
    c[0, 0] = (-3.0, 0.49, 0.5, 300.0)
    assert orc.canvas_to_u8(c)[0, :, 0].tolist() == [0, 0, 1, 255]
    with pytest.raises(ValueError):
        orc.render(model, canvas, 48, 32, np.zeros((57, 32, 48), np.float32), joints, 1, 40)


def test_bench_golden_is_the_oracle(golden_dir):
    """tests/golden/bench_c2.npz (what the -m gpu parity tests compare the benched configuration with) is the oracle's
    own output: frame 3 re-derived live.  Another host CPU may pick another BLAS kernel (different summation order),
    so the live maps are compared at the level two fp32 implementations differ by, and the decisions (noise maps,
    ~800 peaks) may move by a few near-ties."""
    g = np.load(os.path.join(golden_dir, "bench_c2.npz"))
    i = 3
    seed, h, w = [int(v) for v in g["frames"][i]]
    net = orc.Net(orc.COCO_18)
    net.set_weights(synth.make_weights(orc.COCO_18, "he"))
    cnt, joints, peaks, maps = net.process_frame(synth.make_frame(seed, h, w), 368, 656)
    sub = [int(c) for c in g["map_subset"]]
    assert np.abs(maps[:, sub] - g["maps_sub%d" % i]).max() / float(g["maps_absmax%d" % i]) < 1e-5
    assert np.abs(peaks[:, 0, 0] - g["peaks%d" % i][:, 0, 0]).max() <= 2
    assert abs(cnt - int(g["cnt%d" % i])) <= 2
