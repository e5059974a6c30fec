"""BASELINE.json configs at FULL size (parity-test cases, not bench lines):

  C1  MPI 496x368, one 640x480 frame, 1 scale            - whole path vs the oracle
  C2  COCO 656x368, 1 scale, 1280x720 frame              - whole path vs the oracle
  C3  COCO 656x368, 3 scales (1.0/0.85/0.70)             - stride-8 maps + peaks vs the oracle
  C5  COCO 992x736, 4 scales, dense crowd (>= 20 persons) - parse stage, bit-exact, via map injection
      (the conv stack of C5 is the same kernels at a larger M; its oracle forward would take minutes)
plus size-independent properties at full size: batch-position invariance (a frame gives bit-identical results in
any slot of a batch) and determinism across repeated forwards.
"""
import numpy as np
import pytest

from caffe_rtpose_b200 import engine, synth
from oracle import orc
from test_gpu_net import TOL, match_peaks, rel

pytestmark = pytest.mark.gpu


def run_full(model, net_w, net_h, disp_w, disp_h, S, start, gap, prec, thr):
    W = synth.make_weights(model, "he")
    onet = orc.Net(model)
    onet.set_weights(W)
    frame = synth.make_frame(42, disp_h, disp_w)
    ocnt, oj, opk, omaps = onet.process_frame(frame, net_h, net_w, S, start, gap, nms_threshold=thr)
    eng = engine.PoseEngine(model, net_w, net_h, disp_w, disp_h, num_scales=S, start_scale=start, scale_gap=gap, precision=prec)
    eng.set_weights(W)
    eng.nms_layer.SetThreshold(thr)
    eng.forward_frames([frame])
    cnt, joints, peaks = eng.fetch(0)
    maps = eng.fetch_maps(1)
    eng.close()
    return (cnt, joints, peaks, maps), (ocnt, oj, opk, omaps)


def check(got, ref, model, net_w, net_h, S, start, gap, prec, thr):
    cnt, joints, peaks, maps = got
    ocnt, oj, opk, omaps = ref
    assert rel(maps, omaps) < TOL[prec]
    full = orc.imresize(omaps, net_h, net_w, start, gap)
    eps = 20 * TOL[prec] * float(np.abs(omaps).max())
    assert match_peaks(peaks, opk, full, thr, eps) == 0
    return full


def test_c1_mpi_640x480():
    args = (engine.MPI_15, 496, 368, 640, 480, 1, 1.0, 0.3, engine.PREC_BF16X2, 0.8)
    got, ref = run_full(*args)
    check(got, ref, *args[:1], *args[1:3], *args[5:])
    assert got[0] == ref[0]


def test_c2_coco_720p():
    args = (engine.COCO_18, 656, 368, 1280, 720, 1, 1.0, 0.3, engine.PREC_BF16X2, 0.8)
    got, ref = run_full(*args)
    check(got, ref, *args[:1], *args[1:3], *args[5:])
    assert got[0] == ref[0]


def test_c3_three_scales():
    args = (engine.COCO_18, 656, 368, 1280, 720, 3, 1.0, 0.15, engine.PREC_BF16X2, 0.8)
    got, ref = run_full(*args)
    check(got, ref, *args[:1], *args[1:3], *args[5:])


def test_c5_dense_crowd_parse_bit_exact():
    model, net_w, net_h, S = engine.COCO_18, 992, 736, 4
    people = synth.make_people(model, 28, net_w, net_h, seed=21, drop_prob=0.1)
    maps8 = synth.make_maps(model, people, net_w, net_h, num_scales=S, start_scale=1.0, scale_gap=0.15, seed=21)
    eng = engine.PoseEngine(model, net_w, net_h, 1984, 1472, num_scales=S, start_scale=1.0, scale_gap=0.15,
                            precision=engine.PREC_FP32_SIMT)
    eng.forward_maps(maps8)
    cnt, joints, peaks = eng.fetch(0)
    eng.close()
    full = orc.imresize(maps8, net_h, net_w, 1.0, 0.15)
    thr, p = orc.default_params(model)
    opk = orc.nms(full, 18, 64, thr)
    ocnt, oj = orc.connect(model, full, opk, 1984, 1472)
    assert ocnt >= 20 and opk[:, 0, 0].max() <= 64
    assert np.array_equal(peaks, opk) and cnt == ocnt and np.array_equal(joints, oj)


def test_batch_position_invariance_and_determinism():
    """Frames are independent units: the same frame must give bit-identical stride-8 maps, peaks and joints in
    every batch slot and on every repetition (what lets frames shard across GPUs without changing results)."""
    model = engine.COCO_18
    eng = engine.PoseEngine(model, 656, 368, 1280, 720, precision=engine.PREC_BF16X2, max_batch=3)
    eng.set_weights(synth.make_weights(model, "he"))
    eng.nms_layer.SetThreshold(0.8)
    f0, f1 = synth.make_frame(1), synth.make_frame(2)
    eng.forward_frames([f0, f1, f0])
    a = [eng.fetch(i) for i in range(3)]
    ma = eng.fetch_maps(3)
    eng.forward_frames([f1, f0])
    b = [eng.fetch(i) for i in range(2)]
    mb = eng.fetch_maps(2)
    eng.close()
    assert np.array_equal(ma[0], ma[2]) and np.array_equal(ma[0], mb[1]) and np.array_equal(ma[1], mb[0])
    for x, y in ((a[0], a[2]), (a[0], b[1]), (a[1], b[0])):
        assert x[0] == y[0] and np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2])
    assert not np.array_equal(ma[0], ma[1])
