"""TEST INFRASTRUCTURE ONLY - ctypes binding of the CPU oracle (oracle/liboracle.so) and of the
reference-compiled checkers (oracle/_ref/*.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (caffe_rtpose_b200) never does.
"""
import ctypes as C
import glob
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MPI_15, COCO_18 = 0, 1
MAX_PEOPLE = 96

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class ConnectParams(C.Structure):
    _fields_ = [("min_subset_cnt", C.c_int), ("min_subset_score", C.c_float), ("inter_threshold", C.c_float),
                ("inter_min_above", C.c_int), ("clamp_counts", C.c_int)]


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference exists).  Building the checker is not using it."""
    so = os.path.join(HERE, "liboracle.so")
    src = [os.path.join(HERE, f) for f in ("oracle.cpp", "oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    ref = os.environ.get("RTPOSE_REFERENCE", "/root/reference")
    if os.path.isdir(ref) and (force or not os.path.exists(os.path.join(HERE, "_ref", "libref_host.so"))
                               or not os.path.exists(os.path.join(HERE, "_ref", "libref_cpm.so"))):
        subprocess.check_call([sys.executable, os.path.join(HERE, "build_ref.py")], stdout=subprocess.DEVNULL)
    return so


_lib = None


def find_blas():
    """OpenBLAS shipped with the image's wheels (Caffe's BLAS is unpinned: Makefile:369-386)."""
    import site
    roots = site.getsitepackages() + [os.path.dirname(os.path.dirname(np.__file__))]
    for r in roots:
        # scipy's OpenBLAS (LP64, exports scipy_cblas_sgemm) is self-contained; the one bundled with
        # opencv needs its sibling libgfortran/libquadmath preloaded.
        for pat in ("scipy.libs/libscipy_openblas-*.so", "opencv_python_headless.libs/libopenblas*.so*"):
            hits = sorted(glob.glob(os.path.join(r, pat)))
            if hits:
                if "opencv" in pat:
                    for dep in ("libquadmath*", "libgfortran*"):
                        for d in sorted(glob.glob(os.path.join(os.path.dirname(hits[0]), dep))):
                            try:
                                C.CDLL(d, mode=C.RTLD_GLOBAL)
                            except OSError:
                                pass
                return hits[0]
    return None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(os.path.join(HERE, "liboracle.so"))
    L.orc_load_blas.argtypes = [C.c_char_p]
    L.orc_set_threads.argtypes = [C.c_int]
    L.orc_im2col.argtypes = [_f32p] + [C.c_int] * 9 + [_f32p]
    L.orc_conv2d.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p]
    L.orc_maxpool.argtypes = [_f32p] + [C.c_int] * 7 + [_f32p]
    L.orc_relu.argtypes = [_f32p, C.c_size_t]
    L.orc_pooled_dim.argtypes = [C.c_int] * 4
    for f in ("orc_model_limb_seq", "orc_model_map_idx"):
        getattr(L, f).restype = C.POINTER(C.c_int)
    L.orc_model_map_name.restype = C.c_char_p
    L.orc_net_create.restype = C.c_void_p
    L.orc_net_create_stages.restype = C.c_void_p
    L.orc_net_create_stages.argtypes = [C.c_int, C.c_int]
    L.orc_net_destroy.argtypes = [C.c_void_p]
    L.orc_net_num_layers.argtypes = [C.c_void_p]
    L.orc_net_num_convs.argtypes = [C.c_void_p]
    L.orc_net_layer_info.argtypes = [C.c_void_p, C.c_int] + [C.c_char_p] * 4 + [C.POINTER(C.c_int)] * 4
    L.orc_net_conv_info.argtypes = [C.c_void_p, C.c_int, C.c_char_p] + [C.POINTER(C.c_int)] * 3
    L.orc_net_set_weights.argtypes = [C.c_void_p, C.c_char_p, _f32p, _f32p]
    L.orc_net_forward.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, _f32p]
    L.orc_net_forward_blob.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, C.c_char_p, _f32p, C.c_size_t,
                                       C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_net_flops.restype = C.c_double
    L.orc_net_flops.argtypes = [C.c_int, C.c_int, C.c_int]
    L.orc_imresize.argtypes = [_f32p] + [C.c_int] * 6 + [C.c_float, C.c_float, _f32p]
    L.orc_imresize_at.restype = C.c_float
    L.orc_imresize_at.argtypes = [_f32p] + [C.c_int] * 6 + [C.c_float, C.c_float] + [C.c_int] * 3
    L.orc_nms.argtypes = [_f32p] + [C.c_int] * 5 + [C.c_float, _f32p]
    L.orc_default_params.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(ConnectParams)]
    L.orc_connect.argtypes = [C.c_int, _f32p, _f32p] + [C.c_int] * 5 + [C.POINTER(ConnectParams), _f32p, C.c_void_p,
                                                                      C.c_int, C.POINTER(C.c_int)]
    L.orc_resize_area_u8c3.argtypes = [_u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int]
    L.orc_display_scale.restype = C.c_double
    L.orc_display_scale.argtypes = [C.c_int] * 4
    L.orc_warp_affine_cubic_u8c3.argtypes = [_u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, C.c_double]
    L.orc_scale_target.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_preprocess.argtypes = [_u8p] + [C.c_int] * 5 + [C.c_double, C.c_double, _f32p]
    L.orc_json.argtypes = [_f32p, C.c_int, C.c_int, C.c_double, C.c_char_p, C.c_int]
    L.orc_canvas_from_u8.argtypes = [_u8p, C.c_int, C.c_int, _f32p]
    L.orc_canvas_to_u8.argtypes = [_f32p, C.c_int, C.c_int, _u8p]
    L.orc_render.argtypes = [C.c_int, _f32p] + [C.c_int] * 4 + [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int]
    L.orc_render_dispatch.argtypes = [C.c_int, C.c_int, C.c_int, _i32p]
    L.orc_process_frame.argtypes = [C.c_void_p, C.c_int, _u8p] + [C.c_int] * 5 + [C.c_double, C.c_double, C.c_float,
                                                                                C.POINTER(ConnectParams), _f32p,
                                                                                C.c_void_p, C.c_void_p]
    blas = find_blas()
    if blas:
        L.orc_load_blas(blas.encode())
    L.orc_set_threads(os.cpu_count() or 1)
    _lib = L
    return L


# ------------------------------------------------------------------------------------------ helpers
def max_peaks(model):
    return 20 if model == MPI_15 else 64  # prototxt nms_param


def num_parts(model):
    return lib().orc_model_num_parts(model)


def num_maps(model):
    return lib().orc_model_num_maps(model)


def limb_seq(model):
    n = lib().orc_model_num_limbs(model)
    p = lib().orc_model_limb_seq(model)
    return [p[i] for i in range(2 * n)]


def map_idx(model):
    n = lib().orc_model_num_limbs(model)
    p = lib().orc_model_map_idx(model)
    return [p[i] for i in range(2 * n)]


def default_params(model):
    thr = C.c_float()
    p = ConnectParams()
    lib().orc_default_params(model, C.byref(thr), C.byref(p))
    return thr.value, p


def conv2d(x, w, b, pad):
    n, cin, h, ww = x.shape
    cout, _, k, _ = w.shape
    out = np.empty((n, cout, h + 2 * pad - k + 1, ww + 2 * pad - k + 1), np.float32)
    lib().orc_conv2d(np.ascontiguousarray(x, np.float32), n, cin, h, ww, np.ascontiguousarray(w, np.float32),
                     np.ascontiguousarray(b, np.float32), cout, k, pad, out)
    return out


def maxpool(x, k=2, stride=2, pad=0):
    n, c, h, w = x.shape
    L = lib()
    out = np.empty((n, c, L.orc_pooled_dim(h, k, stride, pad), L.orc_pooled_dim(w, k, stride, pad)), np.float32)
    L.orc_maxpool(np.ascontiguousarray(x, np.float32), n, c, h, w, k, stride, pad, out)
    return out


def im2col(im, k, pad, stride=1):
    c, h, w = im.shape
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    col = np.empty((c * k * k, oh * ow), np.float32)
    lib().orc_im2col(np.ascontiguousarray(im, np.float32), c, h, w, k, k, pad, pad, stride, stride, col)
    return col


def imresize(src, th, tw, start_scale, scale_gap):
    num, c, h8, w8 = src.shape
    dst = np.empty((c, th, tw), np.float32)
    lib().orc_imresize(np.ascontiguousarray(src, np.float32), num, c, h8, w8, th, tw, start_scale, scale_gap, dst)
    return dst


def nms(full, n_parts, n_max_peaks, threshold):
    c, h, w = full.shape
    peaks = np.zeros((n_parts, n_max_peaks + 1, 3), np.float32)
    lib().orc_nms(np.ascontiguousarray(full, np.float32), c, h, w, n_parts, n_max_peaks, threshold, peaks)
    return peaks


def connect(model, full, peaks, disp_w, disp_h, params=None, want_subset=False):
    c, net_h, net_w = full.shape
    if params is None:
        _, params = default_params(model)
    P = num_parts(model)
    joints = np.zeros((MAX_PEOPLE, P, 3), np.float32)
    cap = 4096
    subset = np.zeros((cap, P + 3), np.float64)
    rows = C.c_int()
    cnt = lib().orc_connect(model, np.ascontiguousarray(full, np.float32), np.ascontiguousarray(peaks, np.float32),
                            peaks.shape[1] - 1, net_w, net_h, disp_w, disp_h, C.byref(params), joints,
                            subset.ctypes.data_as(C.c_void_p), cap, C.byref(rows))
    if want_subset:
        return cnt, joints[:cnt].copy(), subset[:rows.value].copy()
    return cnt, joints[:cnt].copy()


def resize_area(img, dh, dw):
    sh, sw, _ = img.shape
    out = np.empty((dh, dw, 3), np.uint8)
    rc = lib().orc_resize_area_u8c3(np.ascontiguousarray(img, np.uint8), sh, sw, out, dh, dw)
    if rc:
        raise ValueError("orc_resize_area_u8c3: unsupported (upscale)")
    return out


def display_image(frame, disp_w, disp_h):
    """(display image, frame.scale) of rtpose.cpp:474-487."""
    sh, sw, _ = frame.shape
    s = lib().orc_display_scale(sw, sh, disp_w, disp_h)
    out = np.empty((disp_h, disp_w, 3), np.uint8)
    lib().orc_warp_affine_cubic_u8c3(np.ascontiguousarray(frame, np.uint8), sh, sw, out, disp_h, disp_w, s)
    return out, s


def scale_target(net_w, net_h, start_scale, scale_gap, i):
    tw, th = C.c_int(), C.c_int()
    lib().orc_scale_target(net_w, net_h, start_scale, scale_gap, i, C.byref(tw), C.byref(th))
    return tw.value, th.value


def preprocess(disp, net_h, net_w, num_scales, start_scale, scale_gap):
    dh, dw, _ = disp.shape
    out = np.empty((num_scales, 3, net_h, net_w), np.float32)
    rc = lib().orc_preprocess(np.ascontiguousarray(disp, np.uint8), dh, dw, net_h, net_w, num_scales, start_scale,
                              scale_gap, out)
    if rc:
        raise ValueError("orc_preprocess failed rc=%d" % rc)
    return out


def json_text(joints, n_parts, frame_scale=1.0):
    joints = np.ascontiguousarray(joints, np.float32).reshape(-1, n_parts, 3)
    cap = 64 + joints.shape[0] * (n_parts * 48 + 32)
    buf = C.create_string_buffer(cap)
    n = lib().orc_json(joints if joints.size else np.zeros(1, np.float32), joints.shape[0], n_parts, frame_scale, buf, cap)
    assert n < cap
    return buf.value.decode()


class Net:
    """The deploy graph with Caffe CPU arithmetic."""

    def __init__(self, model, stages=6):
        self.model = model
        self.h = C.c_void_p(lib().orc_net_create_stages(model, stages))

    def __del__(self):
        try:
            lib().orc_net_destroy(self.h)
        except Exception:
            pass

    def convs(self):
        L = lib()
        out = []
        name = C.create_string_buffer(64)
        co, ci, k = C.c_int(), C.c_int(), C.c_int()
        for i in range(L.orc_net_num_convs(self.h)):
            L.orc_net_conv_info(self.h, i, name, C.byref(co), C.byref(ci), C.byref(k))
            out.append((name.value.decode(), co.value, ci.value, k.value))
        return out

    def layers(self):
        L = lib()
        out = []
        bufs = [C.create_string_buffer(256) for _ in range(4)]
        ints = [C.c_int() for _ in range(4)]
        for i in range(L.orc_net_num_layers(self.h)):
            L.orc_net_layer_info(self.h, i, *bufs, *[C.byref(v) for v in ints])
            out.append(dict(name=bufs[0].value.decode(), type=bufs[1].value.decode(),
                            bottom=bufs[2].value.decode().split(","), top=bufs[3].value.decode(),
                            num_output=ints[0].value, kernel_size=ints[1].value, pad=ints[2].value,
                            stride=ints[3].value))
        return out

    def set_weights(self, weights):
        """weights: dict name -> (w[cout,cin,k,k], b[cout])"""
        for name, (w, b) in weights.items():
            rc = lib().orc_net_set_weights(self.h, name.encode(), np.ascontiguousarray(w, np.float32),
                                           np.ascontiguousarray(b, np.float32))
            assert rc == 0, name

    def forward(self, x):
        n, _, h, w = x.shape
        out = np.empty((n, num_maps(self.model), h // 8, w // 8), np.float32)
        rc = lib().orc_net_forward(self.h, np.ascontiguousarray(x, np.float32), n, h, w, out)
        assert rc == 0, rc
        return out

    def forward_blob(self, x, blob, shape):
        n, _, h, w = x.shape
        out = np.empty((n,) + tuple(shape), np.float32)
        bc, bh, bw = C.c_int(), C.c_int(), C.c_int()
        rc = lib().orc_net_forward_blob(self.h, np.ascontiguousarray(x, np.float32), n, h, w, blob.encode(), out,
                                        out.size, C.byref(bc), C.byref(bh), C.byref(bw))
        assert rc == 0 and (bc.value, bh.value, bw.value) == tuple(shape), (rc, bc.value, bh.value, bw.value)
        return out

    def process_frame(self, disp, net_h, net_w, num_scales=1, start_scale=1.0, scale_gap=0.3, nms_threshold=None,
                      params=None):
        thr, p = default_params(self.model)
        if params is not None:
            p = params
        if nms_threshold is not None:
            thr = nms_threshold
        dh, dw, _ = disp.shape
        P, mp = num_parts(self.model), max_peaks(self.model)
        joints = np.zeros((MAX_PEOPLE, P, 3), np.float32)
        peaks = np.zeros((P, mp + 1, 3), np.float32)
        maps8 = np.zeros((num_scales, num_maps(self.model), net_h // 8, net_w // 8), np.float32)
        cnt = lib().orc_process_frame(self.h, self.model, np.ascontiguousarray(disp, np.uint8), dh, dw, net_h, net_w,
                                      num_scales, start_scale, scale_gap, thr, C.byref(p), joints,
                                      peaks.ctypes.data_as(C.c_void_p), maps8.ctypes.data_as(C.c_void_p))
        assert cnt >= 0, cnt
        return cnt, joints[:cnt].copy(), peaks, maps8


def flops(model, h, w):
    return lib().orc_net_flops(model, h, w)


# ------------------------------------------------------------------------------------------ oracle/_ref
_ref_host = None
_ref_cpm = None
_ref_render = None


def ref_host():
    """The reference's own host code compiled from /root/reference (None if not built)."""
    global _ref_host
    if _ref_host is None:
        p = os.path.join(HERE, "_ref", "libref_host.so")
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.ref_connect.argtypes = [C.c_int, _f32p, _f32p] + [C.c_int] * 6 + [C.c_float, C.c_float, C.c_int, _f32p,
                                                                          C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        R.ref_im2col.argtypes = [_f32p] + [C.c_int] * 9 + [_f32p]
        R.ref_model_descriptor.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), _i32p, _i32p, C.c_char_p,
                                           C.c_int]
        if hasattr(R, "ref_maxpool"):   # second translation unit of build_ref.py (absent from libraries built before it existed)
            R.ref_process_and_pad_image.argtypes = [_f32p, _u8p] + [C.c_int] * 5
            R.ref_display_scale.restype = C.c_double
            R.ref_display_scale.argtypes = [C.c_int] * 4
            R.ref_scale_target.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
            R.ref_write_json.argtypes = [C.c_char_p, _f32p, C.c_int, C.c_int, C.c_double]
            R.ref_maxpool.argtypes = [_f32p] + [C.c_int] * 7 + [C.c_void_p, _i32p]
            R.ref_relu.argtypes = [_f32p, _f32p, C.c_int, C.c_float]
        if hasattr(R, "ref_conv_forward"):
            R.ref_load_blas.argtypes = [C.c_char_p]
            R.ref_conv_forward.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_void_p, C.c_int, C.c_int, C.c_int, _f32p]
        if hasattr(R, "ref_render_dispatch"):
            R.ref_render_dispatch.argtypes = [C.c_int, C.c_int, C.c_int, _i32p]
        _ref_host = R
    return _ref_host


def ref_cpm():
    """The reference's own CUDA kernels compiled for sm_100a (needs a GPU to call)."""
    global _ref_cpm
    if _ref_cpm is None:
        p = os.path.join(HERE, "_ref", "libref_cpm.so")
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.ref_imresize_host.argtypes = [_f32p, _f32p] + [C.c_int] * 6 + [C.c_float, C.c_float]
        R.ref_nms_host.argtypes = [_f32p, _f32p] + [C.c_int] * 5 + [C.c_float]
        _ref_cpm = R
    return _ref_cpm


def canvas_from_u8(bgr):
    """process_and_pad_image(normalize=0) of a display-sized image: uint8 HWC -> float planar (rtpose.cpp:239-269)."""
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, _ = bgr.shape
    out = np.zeros((3, h, w), np.float32)
    lib().orc_canvas_from_u8(bgr, h, w, out)
    return out


def canvas_to_u8(canvas):
    """postProcessFrame (rtpose.cpp:1286-1296): float planar canvas -> uint8 HWC."""
    canvas = np.ascontiguousarray(canvas, np.float32)
    _, h, w = canvas.shape
    out = np.zeros((h, w, 3), np.uint8)
    lib().orc_canvas_to_u8(canvas, h, w, out)
    return out


def _poses(poses, num_people, P):
    buf = np.zeros((max(num_people, 1), P, 3), np.float32)
    if num_people:
        buf[:num_people] = np.asarray(poses, np.float32).reshape(-1, P, 3)[:num_people]
    return buf


def render(model, canvas, net_w, net_h, full, poses, num_people, part_to_show=0, googly_eyes=False):
    """CPU restatement of render() + renderFunctions.cu; returns the rendered float canvas (copy)."""
    out = np.ascontiguousarray(canvas, np.float32).copy()
    _, h, w = out.shape
    P = 15 if model == MPI_15 else 18
    fp = None
    if full is not None:
        full = np.ascontiguousarray(full, np.float32)
        fp = full.ctypes.data
    rc = lib().orc_render(model, out, w, h, net_w, net_h, fp, _poses(poses, num_people, P), num_people, part_to_show,
                          1 if googly_eyes else 0)
    if rc != 0:
        raise ValueError("part_to_show %d out of range" % part_to_show)
    return out


def ref_render_lib():
    """The reference's own render kernels compiled for sm_100a (needs a GPU to call)."""
    global _ref_render
    if _ref_render is None:
        p = os.path.join(HERE, "_ref", "libref_render.so")
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.ref_render_host.argtypes = [_f32p] + [C.c_int] * 4 + [C.c_void_p, C.c_int, _f32p] + [C.c_int] * 4
        _ref_render = R
    return _ref_render


def ref_render(model, canvas, net_w, net_h, full, poses, num_people, part_to_show=0, googly_eyes=False):
    R = ref_render_lib()
    out = np.ascontiguousarray(canvas, np.float32).copy()
    _, h, w = out.shape
    P = 15 if model == MPI_15 else 18
    fp, nm = None, 0
    if full is not None:
        full = np.ascontiguousarray(full, np.float32)
        fp, nm = full.ctypes.data, full.shape[0]
    rc = R.ref_render_host(out, w, h, net_w, net_h, fp, nm, _poses(poses, num_people, P), num_people, P, part_to_show,
                           1 if googly_eyes else 0)
    if rc != 0:
        raise RuntimeError("ref_render_host failed")
    return out


def ref_connect(model, full, peaks, disp_w, disp_h, params):
    R = ref_host()
    c, net_h, net_w = full.shape
    P = 15 if model == MPI_15 else 18
    joints = np.zeros((MAX_PEOPLE, P, 3), np.float32)
    cap = 4096
    subset = np.zeros((cap, P + 3), np.float64)
    rows = C.c_int()
    cnt = R.ref_connect(model, np.ascontiguousarray(full, np.float32), np.ascontiguousarray(peaks, np.float32),
                        peaks.shape[1] - 1, net_w, net_h, disp_w, disp_h, params.min_subset_cnt,
                        params.min_subset_score, params.inter_threshold, params.inter_min_above, joints,
                        subset.ctypes.data_as(C.c_void_p), cap, C.byref(rows))
    return cnt, joints[:cnt].copy(), subset[:rows.value].copy()
