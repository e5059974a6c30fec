"""--camera N (cap.open(FLAGS_camera) / cap >> image, rtpose.cpp:401-405, 431) against a capture device.  Neither the build container
nor the GPU boxes have one, so the Video4Linux2 streaming sequence of csrc/camera.cpp runs here against tests/stub/fake_v4l2.c, an
LD_PRELOAD stand-in for /dev/video42 that enforces the driver's state machine (and aborts on a violation), hands out buffers the way
drivers do (fewer than asked, its own frame size, its own format when Motion-JPEG is missing, padded YUYV lines, buffers flagged
V4L2_BUF_FLAG_ERROR) and is "unplugged" after its last frame.  Pixels must equal OpenCV's for the same payloads; rtpose.bin on the
fake camera must push every frame through the worker pipeline in order (stub engine of tests/test_host_pipeline.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "caffe_rtpose_b200")
W, H = 64, 48


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("v4l2") / "fake_v4l2.so")
    r = subprocess.run(["gcc", "-O1", "-g", "-Wall", "-shared", "-fPIC", os.path.join(ROOT, "tests", "stub", "fake_v4l2.c"), "-o", out, "-ldl"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def make_frames(d, fmt, n, stride=None):
    """Payload files for the fake device and the BGR images OpenCV makes of the same payloads."""
    import cv2
    rng = np.random.default_rng(11)
    want = []
    for i in range(n):
        if fmt == "MJPG":
            img = rng.integers(0, 256, (H // 8, W // 8, 3), dtype=np.uint8).repeat(8, 0).repeat(8, 1)
            img[0, 0] = (i, 200, 7)
            ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 95])
            assert ok
            payload = buf.tobytes()
            want.append(cv2.imdecode(buf, cv2.IMREAD_COLOR))
        else:
            yuyv = rng.integers(0, 256, (H, W, 2), dtype=np.uint8)
            yuyv[0, 0] = (16 + i, 128)
            yuyv[0, 1, 1] = 128
            want.append(cv2.cvtColor(yuyv, cv2.COLOR_YUV2BGR_YUYV))
            rows = np.full((H, stride or 2 * W), 0xEE, np.uint8)     # line padding the conversion must skip
            rows[:, :2 * W] = yuyv.reshape(H, 2 * W)
            payload = rows.tobytes()
        (d / ("%03d.bin" % i)).write_bytes(payload)
    return want


GRAB = r"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, %r)
from caffe_rtpose_b200 import engine
L = engine.lib()
h = C.c_void_p()
rc = L.pe_camera_open(42, 1280, 720, C.byref(h))          # asks for 720p, the sensor has its own size
assert rc == 0, L.pe_camera_last_error()
w, hh, cc = C.c_int(), C.c_int(), C.create_string_buffer(5)
L.pe_camera_info(h, C.byref(w), C.byref(hh), cc)
frames = []
while True:
    buf = np.full((hh.value, w.value, 3), 0x5A, np.uint8)
    if L.pe_camera_grab(h, buf.ctypes.data_as(C.c_void_p), buf.size, 2000):
        err = L.pe_camera_last_error().decode()
        break
    frames.append(buf)
L.pe_camera_close(h)
np.savez(sys.argv[1], frames=np.array(frames), size=[w.value, hh.value], fourcc=cc.value.decode(), err=err)
"""


def grab_all(shim, tmp_path, env):
    out = str(tmp_path / "grabbed.npz")
    e = dict(os.environ, LD_PRELOAD=shim, FAKE_V4L2_W=str(W), FAKE_V4L2_H=str(H), **env)
    r = subprocess.run([sys.executable, "-c", GRAB % ROOT, out], capture_output=True, text=True, env=e, timeout=120)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "protocol violation" not in r.stderr
    return np.load(out), r.stderr


@pytest.mark.parametrize("fmt,formats,stride,bad_every", [("MJPG", "MJPG,YUYV", None, 0), ("YUYV", "YUYV", None, 0), ("YUYV", "YUYV", 2 * W + 24, 0),
                                                            ("MJPG", "MJPG,YUYV", None, 3), ("YUYV", "YUYV", None, 4)])
def test_streaming_sequence_against_a_device(shim, tmp_path, fmt, formats, stride, bad_every):
    d = tmp_path / "frames"
    d.mkdir()
    want = make_frames(d, fmt, 10, stride)
    env = {"FAKE_V4L2_FRAMES": str(d), "FAKE_V4L2_FORMATS": formats}
    if stride:
        env["FAKE_V4L2_STRIDE"] = str(stride)
    if bad_every:
        env["FAKE_V4L2_BAD_EVERY"] = str(bad_every)
    got, err = grab_all(shim, tmp_path, env)
    assert list(got["size"]) == [W, H] and str(got["fourcc"]) == fmt          # the driver's size and format, not the requested ones
    assert len(got["frames"]) == 10                                            # flagged buffers are skipped, every real frame arrives, in order
    for a, b in zip(got["frames"], want):
        assert np.array_equal(a, b)
    assert "VIDIOC_DQBUF" in str(got["err"]) and "No such device" in str(got["err"])   # unplugged: reported, not a hang
    delivered = int(err.split("closed after ")[1].split()[0])
    assert delivered == 10 if not bad_every else delivered > 10
    assert "closed while streaming" not in err                                 # VIDIOC_STREAMOFF before close


def test_rtpose_bin_on_the_camera(shim, tmp_path):
    """rtpose.bin without --video / --image_dir opens the camera like the reference; every captured frame goes through a worker and
    the writer in capture order, the run ends when the device disappears (cap >> returns an empty frame)."""
    exe = str(tmp_path / "rtpose_stub")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(PKG, "host", "rtpose.cpp"),
                        os.path.join(ROOT, "tests", "stub", "stub_engine.cpp"), "-o", exe, "-L", PKG, "-lposeengine", "-Wl,-rpath," + PKG, "-pthread"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    d = tmp_path / "frames"
    d.mkdir()
    want = make_frames(d, "YUYV", 25)
    out = tmp_path / "json"
    e = dict(os.environ, LD_PRELOAD=shim, FAKE_V4L2_W=str(W), FAKE_V4L2_H=str(H), FAKE_V4L2_FRAMES=str(d), FAKE_V4L2_FORMATS="YUYV", STUB_FORWARD_MS="1")
    r = subprocess.run([exe, "--model", "COCO", "--caffeproto", "/nonexistent.prototxt", "--random_init", "he", "--camera", "42", "--camera_resolution", "64x48",
                        "--resolution", "%dx%d" % (W, H), "--net_resolution", "32x24", "--write_json", str(out), "--no_frame_drops"],
                       capture_output=True, text=True, env=e, timeout=120)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "Camera 42: 64x48 YUYV" in r.stderr and "protocol violation" not in r.stderr
    assert sorted(os.listdir(out)) == ["frame%06d.json" % i for i in range(25)]
    for i in range(25):
        bodies = json.load(open(out / ("frame%06d.json" % i)))["bodies"]
        b, g = int(want[i][0, 0, 0]), int(want[i][0, 0, 1])
        assert len(bodies) == 1 + b % 3 and bodies[0]["joints"][0] == b + 256 * g     # the stub engine's frame identity


def test_open_errors(shim, tmp_path):
    e = dict(os.environ, LD_PRELOAD=shim, FAKE_V4L2_FORMATS="NV12")
    code = ("import ctypes as C, sys; sys.path.insert(0, %r); from caffe_rtpose_b200 import engine; L = engine.lib(); h = C.c_void_p();"
            "rc = L.pe_camera_open(42, 640, 480, C.byref(h)); print(rc, L.pe_camera_last_error().decode())") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, timeout=60)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.startswith("1 ") and "neither Motion-JPEG nor YUYV" in r.stdout
