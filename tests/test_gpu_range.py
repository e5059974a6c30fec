"""Range management of the fp16-plane parity mode (pe_calibrate / pe_range_status, include/poseengine.h).

The parity mode stores activations as two IEEE fp16 planes.  fp16's exponent range is narrow, so every layer carries a
power-of-two scale (exact; folded into the epilogue's bias / factor).  These tests use nets whose magnitudes are far outside
fp16 without scaling:
  * W-caffe, the prototxt's own filler (gaussian std 0.01, pose_deploy_linevec.prototxt:19-28): every layer shrinks, the stride-8
    maps are ~3e-11 - fp16 planes would flush to zero;
  * He weights x 1.6 per layer: activations grow to ~1e9 - fp16 planes would overflow to inf.
In both cases the uncalibrated engine must REPORT the problem (PE_ERR_RANGE, never silent), and the calibrated engine must match
the oracle at the same 3e-5 (of the map maximum) as everywhere else."""
import numpy as np
import pytest

from caffe_rtpose_b200 import engine, synth
from oracle import orc

pytestmark = pytest.mark.gpu
MODEL, NET_W, NET_H = engine.COCO_18, 160, 96


def oracle_maps(W, frame):
    net = orc.Net(MODEL)
    net.set_weights(W)
    return net.forward(orc.preprocess(frame, NET_H, NET_W, 1, 1.0, 0.3))


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("kind", ["caffe_filler", "growing"])
def test_out_of_range_nets_are_reported_and_calibration_restores_parity(kind):
    if kind == "caffe_filler":
        W = synth.make_weights(MODEL, "caffe")
    else:
        W = {k: (w * np.float32(1.6), b) for k, (w, b) in synth.make_weights(MODEL, "he").items()}
    frame = synth.make_frame(3, 2 * NET_H, 2 * NET_W)
    omaps = oracle_maps(W, frame)
    assert np.isfinite(omaps).all() and np.abs(omaps).max() > 0
    if kind == "caffe_filler":
        assert np.abs(omaps).max() < 1e-8          # ~3e-11: far below fp16's smallest subnormal (6e-8)
    else:
        assert np.abs(omaps).max() > 1e6
    eng = engine.PoseEngine(MODEL, NET_W, NET_H, 2 * NET_W, 2 * NET_H, precision=engine.PREC_F16X2)
    eng.set_weights(W)
    # uncalibrated: the problem is reported, not silent
    eng.forward_frames([frame])
    eng.sync()
    rc, worst, layer = eng.range_status()
    assert rc == 5 and layer, (rc, worst, layer)           # PE_ERR_RANGE, with the offending layer named
    bad = eng.fetch_maps(1)
    assert not np.isfinite(bad).all() or rel(bad, omaps) > 1e-3
    # calibrated: fp32-level parity again, also on the following (CUDA-graph) forwards
    eng.calibrate([frame])
    m0 = eng.fetch_maps(1)
    assert rel(m0, omaps) < 3e-5
    for _ in range(3):
        eng.forward_frames([frame])
    assert np.array_equal(eng.fetch_maps(1), m0)
    rc, worst, layer = eng.range_status()
    assert rc == 0 and 1e-4 < worst < 0.05, (rc, worst, layer)   # stored maxima sit ~1000x below the fp16 limit
    eng.close()


def test_calibration_is_exact_on_a_net_that_did_not_need_it():
    """Scales are powers of two: on the He-init net (activations O(1)) calibrated and uncalibrated engines agree to fp32 rounding
    of the hi/lo split (bit-identical unless a value sits in an fp16 subnormal), and replicas inherit the scales with the weights."""
    W = synth.make_weights(MODEL, "he")
    frame = synth.make_frame(4, 2 * NET_H, 2 * NET_W)
    e0 = engine.PoseEngine(MODEL, NET_W, NET_H, 2 * NET_W, 2 * NET_H, precision=engine.PREC_F16X2)
    e0.set_weights(W)
    e0.forward_frames([frame])
    plain = e0.fetch_maps(1)
    res_plain = e0.fetch(0)
    assert e0.range_status()[0] == 0
    e0.calibrate([frame])
    cal = e0.fetch_maps(1)
    assert rel(cal, plain) < 1e-6
    e1 = engine.PoseEngine(MODEL, NET_W, NET_H, 2 * NET_W, 2 * NET_H, precision=engine.PREC_F16X2)
    engine.share_weights(e0, e1)                      # the scales live in the packed buffer
    e1.forward_frames([frame])
    assert np.array_equal(e1.fetch_maps(1), cal)
    with pytest.raises(engine.PoseEngineError, match="source handle"):
        e1.calibrate([frame])                         # a replica has no fp32 biases to rescale
    cnt, joints, peaks = e1.fetch(0)
    assert cnt == res_plain[0] and np.allclose(peaks, res_plain[2], atol=1e-3)
    e0.close()
    e1.close()
