"""Single-process multi-GPU path (`--num_gpu N`, rtpose.cpp:1459-1549): one handle per GPU, frames sharded by the
shared input queue, weights loaded once and replicated with pe_broadcast_weights (ncclBroadcast).  Needs 2 GPUs; on a
1-GPU box these tests are skipped (the gloo world-2 test in test_sharding.py covers the host logic on CPU)."""
import os
import subprocess

import numpy as np
import pytest

from caffe_rtpose_b200 import engine, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "caffe_rtpose_b200", "rtpose.bin")


def gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.gpu
def test_broadcast_weights_replica_is_bit_identical():
    if gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    model, net_w, net_h, disp_w, disp_h = engine.COCO_18, 160, 96, 320, 192
    W = synth.make_weights(model, "he")
    e0 = engine.PoseEngine(model, net_w, net_h, disp_w, disp_h, precision=engine.PREC_BF16X2, device=0)
    e1 = engine.PoseEngine(model, net_w, net_h, disp_w, disp_h, precision=engine.PREC_BF16X2, device=1)
    e0.set_weights(W)
    engine.broadcast_weights([e0, e1])          # e1 never sees the fp32 weights
    frame = synth.make_frame(5, disp_h, disp_w)
    res = []
    for e in (e0, e1):
        e.forward_frames([frame])
        cnt, joints, peaks = e.fetch(0)
        res.append((cnt, joints.copy(), peaks.copy(), e.fetch_maps(1).copy()))
    assert res[0][0] == res[1][0]
    for a, b in zip(res[0][1:], res[1][1:]):
        assert np.array_equal(a, b)
    # a replica of another configuration is refused
    e2 = engine.PoseEngine(model, net_w, net_h, disp_w, disp_h, precision=engine.PREC_BF16X1, device=1)
    with pytest.raises(engine.PoseEngineError):
        engine.broadcast_weights([e0, e2])
    for e in (e0, e1, e2):
        e.close()


@pytest.mark.gpu
def test_cli_two_gpus_equals_one_gpu(tmp_path):
    if gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    outs = []
    for n in (1, 2):
        out = tmp_path / ("json%d" % n)
        r = subprocess.run([BIN, "--synthetic", "24", "--no_frame_drops", "--random_init", "he", "--model", "COCO", "--resolution", "320x192",
                            "--net_resolution", "160x96", "--write_json", str(out), "--no_display", "--num_gpu", str(n)],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        if n == 2:
            assert "weights broadcast from GPU 0 to 1 replicas" in r.stderr
        outs.append(out)
    names = sorted(os.listdir(outs[0]))
    assert len(names) == 24 and names == sorted(os.listdir(outs[1]))
    for nm in names:
        assert (outs[0] / nm).read_text() == (outs[1] / nm).read_text(), nm


@pytest.mark.gpu
def test_share_weights_same_gpu_is_bit_identical_and_outlives_the_source():
    """Net::ShareTrainedLayersWith (net.cpp:682-706) for two worker handles on one GPU: the second handle uses the first one's
    packed weights without a copy, gives bit-identical results, and keeps working after the source handle is destroyed."""
    model, net_w, net_h, disp_w, disp_h = engine.COCO_18, 160, 96, 320, 192
    e0 = engine.PoseEngine(model, net_w, net_h, disp_w, disp_h, precision=engine.PREC_F16X2)
    e1 = engine.PoseEngine(model, net_w, net_h, disp_w, disp_h, precision=engine.PREC_F16X2, max_batch=2)
    e0.set_weights(synth.make_weights(model, "he"))
    engine.share_weights(e0, e1)
    assert e1.packed_weights() == e0.packed_weights()          # the same device buffer
    frame = synth.make_frame(5, disp_h, disp_w)
    e0.forward_frames([frame])
    want = (e0.fetch(0), e0.fetch_maps(1).copy())
    e0.close()
    e1.forward_frames([frame, frame])
    for i in range(2):
        cnt, joints, peaks = e1.fetch(i)
        assert cnt == want[0][0] and np.array_equal(joints, want[0][1]) and np.array_equal(peaks, want[0][2])
    assert np.array_equal(e1.fetch_maps(2)[:want[1].shape[0]], want[1])
    e2 = engine.PoseEngine(model, net_w, net_h, disp_w, disp_h, precision=engine.PREC_BF16X1)
    with pytest.raises(engine.PoseEngineError):                # another arithmetic mode packs differently
        engine.share_weights(e1, e2)
    e1.close()
    e2.close()
