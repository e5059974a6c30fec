"""rtpose.bin (caffe_rtpose_b200/host/rtpose.cpp): the reference's command line (examples/rtpose/rtpose.cpp:50-72)
over the C ABI.  CPU part: flag surface and error behaviour; GPU part: JSON files identical to the Python path."""
import os
import struct
import subprocess

import numpy as np
import pytest

from caffe_rtpose_b200 import engine, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "caffe_rtpose_b200", "rtpose.bin")

REFERENCE_FLAGS = {  # name -> default (rtpose.cpp:50-72)
    "fullscreen": "false", "part_to_show": "0", "write_frames": "", "no_frame_drops": "false", "write_json": "", "camera": "0",
    "video": "", "image_dir": "", "start_frame": "0", "caffemodel": "model/coco/pose_iter_440000.caffemodel",
    "caffeproto": "model/coco/pose_deploy_linevec.prototxt", "resolution": "1280x720", "net_resolution": "656x368",
    "camera_resolution": "1280x720", "start_device": "0", "num_gpu": "1", "start_scale": "1", "scale_gap": "0.3",
    "num_scales": "1", "no_display": "false", "no_text": "false"}


def run(args, timeout=120):
    return subprocess.run([BIN] + args, capture_output=True, text=True, timeout=timeout)


def test_flag_surface_matches_reference():
    r = run(["--help"])
    assert r.returncode == 0
    for name, dflt in REFERENCE_FLAGS.items():
        assert '--%s (' % name in r.stdout and 'default: "%s"' % dflt in r.stdout, name


def test_flag_errors():
    assert run(["--bogus", "1"]).returncode == 1
    r = run(["--camera", "7"])  # CHECK(cap.open(FLAGS_camera)) (rtpose.cpp:402): no capture device here -> the message, not a hang
    assert r.returncode == 1 and "Couldn't open camera 7 (/dev/video7" in r.stderr
    r = run(["--camera_resolution", "wide"])
    assert r.returncode == 1 and "camera resolution format (wide) invalid" in r.stderr
    r = run(["--video", "/nonexistent/clip.avi", "--model", "COCO"])   # CHECK(cap.open(FLAGS_video)) (rtpose.cpp:406)
    assert r.returncode == 1 and "Couldn't open video file /nonexistent/clip.avi" in r.stderr
    r = run(["--synthetic", "2", "--resolution", "abc", "--model", "COCO"])
    assert r.returncode == 1 and "resolution format" in r.stderr
    r = run(["--synthetic", "2", "--model", "COCO", "--frame_format", "png"])
    assert r.returncode == 1 and "frame_format" in r.stderr


def write_bmp(path, bgr):
    h, w, _ = bgr.shape
    stride = (w * 3 + 3) & ~3
    rows = b"".join(bgr[y].tobytes() + b"\0" * (stride - w * 3) for y in range(h - 1, -1, -1))
    with open(path, "wb") as f:
        f.write(b"BM" + struct.pack("<IHHI", 54 + len(rows), 0, 0, 54))
        f.write(struct.pack("<IiiHHIIiiII", 40, w, h, 1, 24, 0, len(rows), 2835, 2835, 0, 0))
        f.write(rows)


def read_bmp(path):
    d = open(path, "rb").read()
    off, w, h = struct.unpack_from("<I", d, 10)[0], struct.unpack_from("<i", d, 18)[0], struct.unpack_from("<i", d, 22)[0]
    stride = (w * 3 + 3) & ~3
    rows = [np.frombuffer(d, np.uint8, w * 3, off + y * stride).reshape(w, 3) for y in range(h - 1, -1, -1)]
    return np.stack(rows)


def fnv1a(px):
    h = 1469598103934665603
    for b in px.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_image_readers_without_gpu(tmp_path):
    """--image_dir file formats through rtpose.bin's own readers (--probe_image: decode, print size + pixel hash, exit):
    .jpg equals cv::imread (cv2), .png / .bmp / .ppm round-trip."""
    import cv2
    img = synth.make_frame(9, 45, 70)
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])
    (tmp_path / "a.JPG").write_bytes(enc.tobytes())
    want = cv2.imdecode(enc, cv2.IMREAD_COLOR)
    r = run(["--probe_image", str(tmp_path / "a.JPG")])
    assert r.returncode == 0 and r.stdout.split() == ["70x45", "%016x" % fnv1a(want)], (r.stdout, r.stderr)
    (tmp_path / "d.png").write_bytes(cv2.imencode(".png", img)[1].tobytes())
    r = run(["--probe_image", str(tmp_path / "d.png")])
    assert r.returncode == 0 and r.stdout.split() == ["70x45", "%016x" % fnv1a(img)], (r.stdout, r.stderr)
    write_bmp(str(tmp_path / "b.bmp"), img)
    write_ppm(str(tmp_path / "c.ppm"), img)
    for f in ("b.bmp", "c.ppm"):
        r = run(["--probe_image", str(tmp_path / f)])
        assert r.returncode == 0 and r.stdout.split() == ["70x45", "%016x" % fnv1a(img)], f
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    (tmp_path / "p.jpg").write_bytes(enc.tobytes())
    r = run(["--probe_image", str(tmp_path / "p.jpg")])
    assert r.returncode == 0 and r.stdout.split() == ["70x45", "%016x" % fnv1a(cv2.imdecode(enc, cv2.IMREAD_COLOR))]
    ok, enc = cv2.imencode(".jpg", img)
    (tmp_path / "q.jpg").write_bytes(enc.tobytes().replace(b"\xff\xc0", b"\xff\xc9", 1))   # arithmetic-coded: not handled
    r = run(["--probe_image", str(tmp_path / "q.jpg")])
    assert r.returncode == 1 and "not handled" in r.stderr


def test_multi_producer_decode_stage_without_gpu(tmp_path):
    """--num_producers N (extension): N decoder threads feed the queue, every file index exactly once, an undecodable file
    becomes a dropped index; --decode_bench runs this stage alone."""
    import cv2
    for i in range(12):
        cv2.imwrite(str(tmp_path / ("f%02d.jpg" % i)), synth.make_frame(i, 90, 160), [cv2.IMWRITE_JPEG_QUALITY, 85])
    (tmp_path / "f05.jpg").write_bytes(b"\xff\xd8 not a jpeg")
    for n in (1, 4):
        r = run(["--image_dir", str(tmp_path), "--decode_bench", "--num_producers", str(n), "--model", "COCO", "--resolution", "160x90"])
        assert r.returncode == 0, r.stderr
        out = r.stdout.strip().splitlines()[-1]
        assert out.startswith("decoded 11 frames") and "indices_unique 1" in out and ("with %d producer" % n) in out, out
        assert ("dropped 1" in out) == (n > 1)   # the single-thread path skips the file without consuming an index


def write_mjpeg_avi(path, frames, fps=25.0):
    import cv2
    h, w, _ = frames[0].shape
    wr = cv2.VideoWriter(path, cv2.CAP_OPENCV_MJPEG, cv2.VideoWriter_fourcc(*"MJPG"), fps, (w, h))
    assert wr.isOpened()
    for f in frames:
        wr.write(f)
    wr.release()


def test_video_source_decode_stage_without_gpu(tmp_path):
    """--video (rtpose.cpp:394-411): a Motion-JPEG .avi goes through the producer stage; --resolution -1x-1 takes the video's size
    (:1677-1682); --video_realtime false lets several decoder threads read one file; the reference's pacing holds the file's frame
    rate; --start_frame skips."""
    path = str(tmp_path / "clip.avi")
    write_mjpeg_avi(path, [synth.make_frame(i, 90, 160) for i in range(10)], fps=50.0)
    for n in (1, 3):
        r = run(["--video", path, "--decode_bench", "--novideo_realtime", "--num_producers", str(n), "--model", "COCO", "--resolution", "-1x-1",
                 "--write_json", str(tmp_path / "j")])
        assert r.returncode == 0, r.stderr
        assert "Video %s: 160x90, 10 frames, 50.000 fps, MJPG" % path in r.stderr and "Display resolution: 160x90" in r.stderr
        out = r.stdout.strip().splitlines()[-1]
        assert out.startswith("decoded 10 frames") and "indices_unique 1" in out and ("with %d producer" % n) in out, out
    r = run(["--video", path, "--decode_bench", "--model", "COCO", "--resolution", "160x90", "--start_frame", "4"])   # paced: 6 frames at 50 fps
    assert r.returncode == 0, r.stderr
    out = r.stdout.strip().splitlines()[-1]
    assert out.startswith("decoded 6 frames") and "with 1 producer" in out
    secs = float(out.split(" in ")[1].split(" s")[0])
    assert 0.09 < secs < 1.0, out          # 5 frame intervals of 20 ms
    # handleKey's video keys (rtpose.cpp:1572-1593) from stdin: 'l' jumps 30 frames ahead, so a 120-frame clip ends ~30 frames early
    long_clip = str(tmp_path / "long.avi")
    write_mjpeg_avi(long_clip, [synth.make_frame(i % 3, 48, 64) for i in range(120)], fps=200.0)
    p = subprocess.run([BIN, "--video", long_clip, "--decode_bench", "--model", "COCO", "--resolution", "64x48", "--keys_from_stdin"],
                       input="l", capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    n = int(p.stdout.strip().splitlines()[-1].split()[1])
    assert "Seek to frame" in p.stderr and 85 <= n <= 95, (n, p.stderr[-300:])


def write_ppm(path, bgr):
    h, w, _ = bgr.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(bgr[:, :, ::-1]).tobytes())


@pytest.mark.gpu
def test_cli_json_equals_python_path(tmp_path):
    model, net_w, net_h, disp_w, disp_h = engine.COCO_18, 160, 96, 320, 192
    W = synth.make_weights(model, "he")
    cm = str(tmp_path / "pose.caffemodel")
    engine.write_caffemodel(cm, W, synth.conv_table(model))
    proto = tmp_path / "deploy.prototxt"
    # the deploy prototxt itself (rebuilt from the committed layer table of model/coco/pose_deploy_linevec.prototxt): rtpose.bin
    # builds the net from it like `new Net<float>(proto, TEST)`
    import json
    proto.write_text(synth.netspec_to_prototxt(json.load(open(os.path.join(ROOT, "tests", "golden", "netspec_coco.json")))))
    img_dir = tmp_path / "frames"
    img_dir.mkdir()
    frames = [synth.make_frame(20 + i, disp_h, disp_w) for i in range(5)]
    for i, f in enumerate(frames):
        (write_bmp if i % 2 == 0 else write_ppm)(str(img_dir / ("img%03d.%s" % (i, "bmp" if i % 2 == 0 else "ppm"))), f)
    out = tmp_path / "json"
    r = run(["--image_dir", str(img_dir), "--caffemodel", cm, "--caffeproto", str(proto), "--resolution", "%dx%d" % (disp_w, disp_h),
             "--net_resolution", "%dx%d" % (net_w, net_h), "--write_json", str(out), "--no_display", "--no_frame_drops", "--nocalibrate_range", "--num_gpu", "1",
             "--write_frames", str(tmp_path / "rendered"), "--part_to_show", "2", "--no_text"], timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    eng = engine.PoseEngine(model, net_w, net_h, disp_w, disp_h, precision=engine.PREC_BF16X2)
    eng.set_weights(W)
    for i, f in enumerate(frames):
        eng.forward_frames([f])
        cnt, joints, _ = eng.fetch(0)
        got = (out / ("img%03d.json" % i)).read_text()
        assert got == eng.json(joints, 1.0)
        # --write_frames: the rendered display image (heat map of part 1 here) as quality-98 JPEG, like the reference
        assert (tmp_path / "rendered" / ("img%03d.jpg" % i)).read_bytes() == engine.encode_jpeg(eng.render(0, 2), 98)
    eng.close()
    # frames whose size differs from --resolution go through the GPU warpAffine (rtpose.cpp:474-487); JSON carries 1/scale
    big_dir = tmp_path / "big"
    big_dir.mkdir()
    big = synth.make_frame(77, 270, 480)
    write_bmp(str(big_dir / "big.bmp"), big)
    out2 = tmp_path / "json2"
    r = run(["--image_dir", str(big_dir), "--caffemodel", cm, "--caffeproto", str(proto), "--resolution", "320x192",
             "--net_resolution", "160x96", "--write_json", str(out2), "--no_display", "--no_frame_drops", "--nocalibrate_range"], timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    eng = engine.PoseEngine(model, net_w, net_h, disp_w, disp_h, precision=engine.PREC_BF16X2)
    eng.set_weights(W)
    sc = eng.forward_camera_frames([big])
    cnt, joints, _ = eng.fetch(0)
    assert abs(sc - 320 / 480.0) < 1e-12
    assert (out2 / "big.json").read_text() == eng.json(joints, sc)
    eng.close()
    # --video: the frames of a Motion-JPEG .avi (other size than --resolution: GPU warpAffine) give the JSON files the same pixels give
    # through the Python path; files are named after the video frame number (rtpose.cpp:1399-1402)
    clip = str(tmp_path / "clip.avi")
    vframes = [synth.make_frame(40 + i, 270, 480) for i in range(4)]
    write_mjpeg_avi(clip, vframes)
    out3 = tmp_path / "json3"
    r = run(["--video", clip, "--novideo_realtime", "--caffemodel", cm, "--caffeproto", str(proto), "--resolution", "320x192",
             "--net_resolution", "160x96", "--write_json", str(out3), "--no_display", "--no_frame_drops", "--nocalibrate_range"], timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    cap = engine.VideoCapture(clip)
    eng = engine.PoseEngine(model, net_w, net_h, disp_w, disp_h, precision=engine.PREC_BF16X2)
    eng.set_weights(W)
    for i in range(4):
        ok, fr = cap.read()
        assert ok
        sc = eng.forward_camera_frames([fr])
        cnt, joints, _ = eng.fetch(0)
        assert (out3 / ("frame%06d.json" % i)).read_text() == eng.json(joints, sc)
    eng.close()
    # lossless frames on request
    r = run(["--image_dir", str(big_dir), "--caffemodel", cm, "--caffeproto", str(proto), "--resolution", "320x192",
             "--net_resolution", "160x96", "--no_display", "--no_frame_drops", "--nocalibrate_range", "--write_frames", str(tmp_path / "bmp"), "--frame_format", "bmp", "--no_text"], timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    eng = engine.PoseEngine(model, net_w, net_h, disp_w, disp_h, precision=engine.PREC_BF16X2)
    eng.set_weights(W)
    eng.forward_camera_frames([big])
    plain = read_bmp(str(tmp_path / "bmp" / "big.bmp"))
    assert np.array_equal(plain, eng.render(0, 0))
    eng.close()
    # without --no_text the fps / people-count overlays of displayFrame (rtpose.cpp:1317-1353) are drawn: only the top band changes
    r = run(["--image_dir", str(big_dir), "--caffemodel", cm, "--caffeproto", str(proto), "--resolution", "320x192",
             "--net_resolution", "160x96", "--no_display", "--no_frame_drops", "--nocalibrate_range", "--write_frames", str(tmp_path / "txt"), "--frame_format", "bmp"], timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    texted = read_bmp(str(tmp_path / "txt" / "big.bmp"))
    assert np.array_equal(texted[60:], plain[60:]) and (texted[:60] != plain[:60]).any()
    # --resolution -1x-1 takes the size from the first image (rtpose.cpp:1683-1686); missing model file is an error
    r = run(["--image_dir", str(img_dir), "--caffemodel", cm, "--caffeproto", str(proto), "--resolution", "-1x-1",
             "--net_resolution", "%dx%d" % (net_w, net_h), "--no_display"], timeout=300)
    assert r.returncode == 0 and "from first image: 320x192" in r.stderr
    r = run(["--image_dir", str(img_dir), "--caffemodel", str(tmp_path / "none.caffemodel"), "--caffeproto", str(proto),
             "--resolution", "320x192", "--net_resolution", "160x96"], timeout=300)
    assert r.returncode == 1 and "cannot load" in r.stderr


@pytest.mark.gpu
def test_frame_drop_policy_latency_line_and_runtime_keys(tmp_path):
    """processFrame drops frames that waited more than 0.1 s for a GPU unless --no_frame_drops (rtpose.cpp:1107-1124); the
    30-frame status line carries the reference's stage names (:1421-1441); handleKey's threshold keys (:1617-1651) arrive on
    stdin.  A producer that is far faster than a 1-frame-per-forward worker on a tiny net shows all three."""
    common = ["--synthetic", "400", "--random_init", "he", "--model", "COCO", "--resolution", "320x192", "--net_resolution", "160x96",
              "--no_display", "--batch", "1", "--num_producers", "4"]
    out = tmp_path / "json"
    r = run(common + ["--write_json", str(out), "--no_frame_drops"], timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(os.listdir(out)) == 400 and "0 dropped" in r.stderr
    assert "Latency" in r.stderr and "QueueA" in r.stderr and "Buffered" in r.stderr and "FPS =" in r.stderr
    out2 = tmp_path / "json_drop"
    r = run(common + ["--write_json", str(out2)], timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    n = len(os.listdir(out2))
    dropped = int(r.stderr.split(" dropped")[0].split()[-1])
    assert n + dropped == 400                      # every frame is either written or accounted as dropped, order preserved
    # keys: two '=' raise the NMS threshold by 0.01, ']' raises connect_inter_threshold
    p = subprocess.run([BIN] + common + ["--keys_from_stdin", "--no_frame_drops"], input="==]\n", capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "nms_threshold: 0.06" in p.stderr and "connect_inter_threshold: 0.055" in p.stderr
