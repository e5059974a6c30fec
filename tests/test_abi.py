"""CPU tests of the drop-in boundary: libposeengine.so loads, exports every symbol include/poseengine.h
declares, fails loudly without a GPU (no CPU fallback), and its host-only entry points (JSON writer, model
descriptor tables) match the oracle and the reference's own modelDescriptorFactory.cpp."""
import ctypes as C
import os
import struct
import re
import subprocess

import numpy as np
import pytest

from caffe_rtpose_b200 import engine, synth
from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "poseengine.h")).read()
    declared = sorted(set(re.findall(r"\b(pe_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 35
    L = C.CDLL(engine.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), "missing export %s" % name
    assert sorted(engine.ABI_SYMBOLS) == declared


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.PoseEngineError, match="no CPU fallback"):
        engine.PoseEngine()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "caffe_rtpose_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("(see oracle.cpp)", "").replace("the oracle", "").replace("to the oracle", ""), f


@pytest.mark.parametrize("model", [engine.MPI_15, engine.COCO_18])
def test_model_descriptor_api(model):
    md = engine.ModelDescriptorFactory.createModelDescriptor(model)
    assert md.get_number_parts() == orc.num_parts(model)
    assert md.number_limb_sequence() == len(orc.limb_seq(model)) // 2
    assert md.get_limb_sequence() == orc.limb_seq(model)
    assert md.get_map_idx() == orc.map_idx(model)
    n = md.get_number_parts() + 1 + 2 * md.number_limb_sequence()
    assert [md.get_part_name(i) for i in range(n)] == [orc.lib().orc_model_map_name(model, i).decode() for i in range(n)]
    with pytest.raises(IndexError):
        md.get_part_name(n)
    with pytest.raises(RuntimeError):
        engine.ModelDescriptorFactory.createModelDescriptor(7)


def test_json_writer_matches_oracle():
    rng = np.random.default_rng(0)
    j = (rng.uniform(0, 1300, (3, 18, 3))).astype(np.float32)
    j[1, 4] = 0
    for scale in (1.0, 0.5, 1.7777):
        assert engine.write_json(j, 18, scale) == orc.json_text(j, 18, scale)
    assert engine.write_json(np.zeros((0, 15, 3), np.float32), 15) == orc.json_text(np.zeros((0, 15, 3), np.float32), 15)


def test_weight_file_roundtrip(tmp_path):
    w = synth.make_weights(engine.MPI_15, "caffe")
    p = str(tmp_path / "w.rtpw")
    engine.write_weights_file(p, w, synth.conv_table(engine.MPI_15))
    raw = open(p, "rb").read()
    assert raw[:4] == b"RTPW" and len(raw) > 4 * 51000000


@pytest.mark.parametrize("v1,legacy_dims", [(False, False), (True, True), (False, True)])
def test_caffemodel_wire_reader(tmp_path, v1, legacy_dims):
    """pe_caffemodel_*: binary NetParameter walker (caffe.proto:10-22, 92-95, 311-329, 1272-1276), host only."""
    table = [("conv1_1", 4, 3, 3), ("Mconv7_stage6_L2", 5, 7, 1)]
    rng = np.random.default_rng(1)
    w = {n: (rng.standard_normal((co, ci, k, k)).astype(np.float32), rng.standard_normal(co).astype(np.float32))
         for n, co, ci, k in table}
    p = str(tmp_path / "m.caffemodel")
    engine.write_caffemodel(p, w, table, legacy_v1=v1, legacy_dims=legacy_dims)
    layers = [l for l in engine.read_caffemodel(p) if l[2]]
    assert [l[0] for l in layers] == [t[0] for t in table]
    for (name, typ, blobs), (n, co, ci, k) in zip(layers, table):
        assert typ == ("V1:4" if v1 else "Convolution")
        assert np.array_equal(blobs[0][0], w[n][0].ravel()) and np.array_equal(blobs[1][0], w[n][1])
        assert blobs[0][1] == (co, ci, k, k)
        assert blobs[1][1] == ((1, 1, 1, co) if legacy_dims else (co,))
    with pytest.raises(engine.PoseEngineError):
        engine.read_caffemodel(str(tmp_path / "missing.caffemodel"))
    bad = tmp_path / "bad.caffemodel"
    bad.write_bytes(b"\x92\x06\xff\xff\xff\xff\x0f garbage")
    with pytest.raises(engine.PoseEngineError):
        engine.read_caffemodel(str(bad))


@pytest.mark.parametrize("fname,v1", [("caffemodel_v2.caffemodel", False), ("caffemodel_v1.caffemodel", True)])
def test_caffemodel_written_by_google_protobuf(fname, v1):
    """Files NOT produced by this repo's writer: tools/gen_caffemodel_fixture.py serialised them with Google's protobuf runtime from
    descriptors restating caffe.proto - `layer` + BlobShape + packed data + skipped fields (bottom/top/param/phase/convolution_param,
    diff arrays, blob-less layers), and legacy `layers` + num/channels/height/width + UNPACKED floats + blobs_lr/weight_decay."""
    gold = os.path.join(ROOT, "tests", "golden")
    want = np.load(os.path.join(gold, "caffemodel_fixture.npz"))
    layers = engine.read_caffemodel(os.path.join(gold, fname))
    conv = [l for l in layers if l[2]]
    assert [l[0] for l in conv] == ["conv1_1", "conv5_5_CPM_L2", "Mconv7_stage6_L1"]
    assert len(layers) == (3 if v1 else 6)                       # V2 file also holds three blob-less ReLU layers
    for name, typ, blobs in conv:
        assert typ == ("V1:4" if v1 else "Convolution")
        w, b = want[name + "_w"], want[name + "_b"]
        assert len(blobs) == 2
        assert np.array_equal(blobs[0][0], w.ravel()) and blobs[0][1] == w.shape
        assert np.array_equal(blobs[1][0], b) and blobs[1][1] == ((1, 1, 1, len(b)) if v1 else (len(b),))


def test_header_is_plain_c_and_links(tmp_path):
    """include/poseengine.h must be consumable from C (the ABI a cgo/JNI/ctypes/C++ caller binds) and every call
    used by the reference-side stub of INTEGRATION.md must link against libposeengine.so."""
    import subprocess
    src = tmp_path / "abi_check.c"
    src.write_text(r"""
#include "poseengine.h"
#include <stdio.h>
int main(void) {
    pe_config c = {0};
    pe_engine* e = 0;
    c.model = PE_MODEL_COCO_18; c.net_w = 656; c.net_h = 368; c.disp_w = 1280; c.disp_h = 720;
    c.num_scales = 1; c.start_scale = 1.0; c.scale_gap = 0.3; c.max_batch = 1; c.precision = PE_PREC_BF16X2;
    int rc = pe_create(&c, &e);
    printf("%d|%s|%d|%d\n", rc, pe_last_error(0), pe_model_num_parts(PE_MODEL_MPI_15), pe_model_limb_sequence(PE_MODEL_COCO_18)[3]);
    float j[3] = {1.5f, 2.5f, 0.5f};
    char buf[256];
    pe_write_json(j, 0, 18, 1.0, buf, 256);
    printf("%s", buf);
    if (e) pe_destroy(e);
    return 0;
}
""")
    exe = tmp_path / "abi_check"
    libdir = os.path.dirname(engine.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lposeengine", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120).stdout
    first = out.split("\n")[0].split("|")
    assert first[2] == "15" and first[3] == "5"
    import torch
    if not torch.cuda.is_available():
        assert first[0] == "2" and "no CPU fallback" in first[1]
    assert '"bodies":[' in out


def test_cpp_header_shims_compile(tmp_path):
    """include/rtpose/*.h and include/caffe/cpm/layers/*.hpp: the reference's class names and methods over the ABI."""
    import subprocess
    src = tmp_path / "shim_check.cpp"
    src.write_text(r"""
#include "rtpose/modelDescriptorFactory.h"
#include "caffe/cpm/layers/nms_layer.hpp"
#include "caffe/cpm/layers/imresize_layer.hpp"
#include "rtpose/renderFunctions.h"
#include <cstdio>
// the three render entry points keep the reference's signatures (renderFunctions.h:8-17): taking their addresses with the
// reference's function types must compile
typedef void (*mpi_fn)(float*, int, int, int, int, float*, int, float*, float*, std::vector<int>, int);
typedef void (*coco_fn)(float*, int, int, int, int, float*, int, float*, float*, std::vector<int>, int, bool);
typedef void (*aff_fn)(float*, int, int, int, int, float*, int, float*, float*, std::vector<int>, int, int);
static mpi_fn f0 = &render_mpi_parts; static coco_fn f1 = &render_coco_parts; static aff_fn f2 = &render_coco_aff;
int main() {
    if (!f0 || !f1 || !f2 || RENDER_MAX_PEOPLE != 96) return 1;
    std::unique_ptr<ModelDescriptor> md;
    ModelDescriptorFactory::createModelDescriptor(ModelDescriptorFactory::Type::COCO_18, md);
    printf("%d %d %s %s\n", md->get_number_parts(), md->number_limb_sequence(), md->get_part_name(19).c_str(),
           md->get_part_name(md->get_map_idx()[0]).c_str());
    try { md->get_part_name(99); } catch (const std::out_of_range&) { printf("oor\n"); }
    try { ModelDescriptor bad({{0, "a"}, {1, "b"}}, {0, 1}, {2}); } catch (const std::runtime_error&) { printf("rte\n"); }
    caffe::NmsLayer<float>* nms = nullptr; caffe::ImResizeLayer<float>* rs = nullptr; (void)nms; (void)rs;
    return 0;
}
""")
    exe = tmp_path / "shim_check"
    libdir = os.path.dirname(engine.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir,
                           "-lposeengine", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120).stdout.split("\n")
    assert out[0] == "18 19 %s %s" % (orc.lib().orc_model_map_name(orc.COCO_18, 19).decode(), orc.lib().orc_model_map_name(orc.COCO_18, 31).decode())
    assert out[1] == "oor" and out[2] == "rte"


def test_jpeg_encoder_decodes_like_libjpeg_at_same_quality():
    """pe_encode_jpeg (--write_frames, rtpose.cpp:1363-1380 uses cv::imwrite quality 98): a standard JFIF stream that
    OpenCV's libjpeg decodes, with the same quality/size trade-off as cv2.imencode at the same setting."""
    import cv2
    from caffe_rtpose_b200 import synth

    def psnr(a, b):
        d = a.astype(np.float64) - b.astype(np.float64)
        return 10 * np.log10(255.0 ** 2 / max(np.mean(d * d), 1e-12))

    for h, w in [(192, 320), (100, 75), (17, 33), (8, 8), (1, 1)]:
        img = synth.make_frame(3, h, w)
        if min(h, w) > 8:
            img = cv2.GaussianBlur(img, (0, 0), 3)   # camera-like content; pure noise is chroma-subsampled away by any JPEG
        for q in (98, 75, 30):
            data = engine.encode_jpeg(img, q)
            assert data[:2] == b"\xff\xd8" and data[-2:] == b"\xff\xd9"
            dec = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
            assert dec is not None and dec.shape == img.shape
            ok, ref = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])
            rdec = cv2.imdecode(ref, cv2.IMREAD_COLOR)
            assert psnr(dec, img) > psnr(rdec, img) - 0.5, (h, w, q)
            assert len(data) < 1.25 * len(ref) + 64, (h, w, q)
    # extreme block contrast at quality 100 stays inside the baseline coefficient range
    chk = (np.indices((16, 16)).sum(0) % 2 * 255).astype(np.uint8)[:, :, None].repeat(3, 2)
    dec = cv2.imdecode(np.frombuffer(engine.encode_jpeg(chk, 100), np.uint8), cv2.IMREAD_COLOR)
    assert psnr(dec, chk) > 30
    with pytest.raises(engine.PoseEngineError):
        engine.encode_jpeg(np.zeros((0, 4, 3), np.uint8))


def _jpeg_variants(cv2, q):
    base = [cv2.IMWRITE_JPEG_QUALITY, q]
    prog = base + [cv2.IMWRITE_JPEG_PROGRESSIVE, 1]
    v = [("420", base), ("optimised-huffman", base + [cv2.IMWRITE_JPEG_OPTIMIZE, 1]), ("restart-3", base + [cv2.IMWRITE_JPEG_RST_INTERVAL, 3]),
         ("progressive", prog), ("progressive+restart", prog + [cv2.IMWRITE_JPEG_RST_INTERVAL, 2])]
    if hasattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR"):
        for name, f in (("444", cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444), ("422", cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422)):
            v += [(name, base + [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, f]), ("progressive-" + name, prog + [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, f])]
    return v


def test_jpeg_decoder_equals_cv_imread_bit_for_bit():
    """pe_decode_jpeg feeds --image_dir where the reference calls cv::imread (rtpose.cpp:302-391): same pixels as OpenCV's
    libjpeg for baseline AND progressive files - sizes that are not multiples of the MCU, 4:2:0 / 4:2:2 / 4:4:4, grey,
    optimised Huffman tables, restart intervals, and the stream of our own encoder."""
    import cv2
    from caffe_rtpose_b200 import synth

    for h, w in [(192, 320), (100, 75), (17, 33), (8, 8), (1, 1), (16, 16), (31, 47), (3, 5), (2, 4), (5, 3)]:
        noise = synth.make_frame(3, h, w)
        for img in (noise, cv2.GaussianBlur(noise, (0, 0), 2) if min(h, w) > 8 else noise):
            for q in (98, 75, 20):
                for name, params in _jpeg_variants(cv2, q):
                    ok, enc = cv2.imencode(".jpg", img, params)
                    assert ok
                    assert np.array_equal(engine.decode_jpeg(enc.tobytes()), cv2.imdecode(enc, cv2.IMREAD_COLOR)), (h, w, q, name)
        for params in ([cv2.IMWRITE_JPEG_QUALITY, 90], [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_PROGRESSIVE, 1]):
            ok, enc = cv2.imencode(".jpg", noise[:, :, 0], params)
            assert np.array_equal(engine.decode_jpeg(enc.tobytes()), cv2.imdecode(enc, cv2.IMREAD_COLOR))
        own = engine.encode_jpeg(noise, 98)
        assert np.array_equal(engine.decode_jpeg(own), cv2.imdecode(np.frombuffer(own, np.uint8), cv2.IMREAD_COLOR))
    ok, enc = cv2.imencode(".jpg", synth.make_frame(1, 64, 64))
    arith = enc.tobytes().replace(b"\xff\xc0", b"\xff\xc9", 1)   # SOF9 = arithmetic coding: a variant that is not handled
    with pytest.raises(engine.PoseEngineError, match="unsupported"):
        engine.decode_jpeg(arith)
    with pytest.raises(engine.PoseEngineError, match="not a JPEG"):
        engine.decode_jpeg(b"BM" + bytes(100))
    with pytest.raises(engine.PoseEngineError):
        engine.decode_jpeg(enc.tobytes()[:40])


def _decode_route(data, fast):
    """(return code, pixels) of pe_decode_jpeg on one of its two routes (PE_JPEG_FAST is read per call)."""
    os.environ["PE_JPEG_FAST"] = "1" if fast else "0"
    try:
        try:
            return 0, engine.decode_jpeg(data)
        except engine.PoseEngineError as ex:
            return 1, str(ex)
    finally:
        del os.environ["PE_JPEG_FAST"]


def test_jpeg_fast_route_equals_general_route_also_on_corrupt_streams():
    """The sequential fast route of jpeg_dec.cpp (8-byte refills, combined code+value lookahead, AVX2 islow IDCT with its range
    guard, fused row-wise output) is only allowed to be faster: same pixels and same verdict as the general route on valid
    files of every sampling / restart / Huffman flavour and on byte-mutated and truncated streams (where the 32-bit IDCT guard
    and the stop rules of the entropy decoder matter)."""
    import cv2
    from caffe_rtpose_b200 import synth

    rng = np.random.default_rng(7)
    streams = []
    for h, w in [(96, 160), (37, 53), (16, 16), (9, 250)]:
        img = synth.make_frame(5, h, w)
        flat = np.full((h, w, 3), 255, np.uint8)
        flat[::2] = 0                                   # extreme contrast: large coefficients
        for pic in (img, flat, rng.integers(0, 256, (h, w, 3), dtype=np.uint8)):
            for q in (100, 90, 30, 3):
                for name, params in _jpeg_variants(cv2, q):
                    if "progressive" not in name:
                        streams.append(cv2.imencode(".jpg", pic, params)[1].tobytes())
            streams.append(cv2.imencode(".jpg", pic[:, :, 1], [cv2.IMWRITE_JPEG_QUALITY, 85])[1].tobytes())
            streams.append(bytes(engine.encode_jpeg(pic, 98)))
    n_ok = n_bad = 0
    for data in streams:
        rc_f, px_f = _decode_route(data, True)
        rc_g, px_g = _decode_route(data, False)
        assert rc_f == rc_g == 0 and np.array_equal(px_f, px_g)
    for data in streams[::3]:
        sos = data.index(b"\xff\xda")
        for trial in range(12):
            buf = bytearray(data)
            kind = trial % 4
            if kind == 0:      # flip bytes inside the entropy-coded data
                for _ in range(int(rng.integers(1, 4))):
                    buf[int(rng.integers(sos + 12, len(buf) - 2))] = int(rng.integers(0, 256))
            elif kind == 1:    # flip bytes in the tables / headers
                buf[int(rng.integers(4, sos))] = int(rng.integers(0, 256))
            elif kind == 2:    # truncate
                buf = buf[:int(rng.integers(sos, len(buf)))]
            else:              # inflate a quantiser entry and a few data bytes: drives blocks past the 32-bit IDCT guard
                dqt = data.index(b"\xff\xdb")
                buf[dqt + 5 + int(rng.integers(0, 8))] = 255
                buf[int(rng.integers(sos + 12, len(buf) - 2))] = int(rng.integers(0, 256))
            rc_f, px_f = _decode_route(bytes(buf), True)
            rc_g, px_g = _decode_route(bytes(buf), False)
            assert rc_f == rc_g, (trial, rc_f, rc_g)
            if rc_f == 0:
                assert np.array_equal(px_f, px_g), trial
                n_ok += 1
            else:
                n_bad += 1
    assert n_ok > 100 and n_bad > 5   # the mutations are neither all harmless nor all fatal


def test_jpeg_idct_scalar_fallback_equals_avx2_form(tmp_path):
    """The AVX2 IDCT hands a block to the 64-bit scalar form when a value leaves the 32-bit-safe range.  A build whose bound is
    64 (almost every block falls back) must produce the library's pixels."""
    import ctypes as C

    import cv2
    from caffe_rtpose_b200 import synth

    so = str(tmp_path / "libjpegdec_b64.so")
    src = os.path.join(ROOT, "caffe_rtpose_b200", "csrc", "jpeg_dec.cpp")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DPE_IDCT32_BOUND=64", "-I", os.path.join(ROOT, "include"), src, "-o", so],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    alt = C.CDLL(so)
    alt.pe_decode_jpeg.argtypes = [C.c_char_p, C.c_longlong, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_longlong]
    for h, w, q in [(120, 200, 95), (64, 64, 40), (33, 70, 100)]:
        data = cv2.imencode(".jpg", synth.make_frame(2, h, w), [cv2.IMWRITE_JPEG_QUALITY, q])[1].tobytes()
        ww, hh = C.c_int(), C.c_int()
        out = np.zeros((h, w, 3), np.uint8)
        assert alt.pe_decode_jpeg(data, len(data), C.byref(ww), C.byref(hh), out.ctypes.data, out.size) == 0
        assert np.array_equal(out, engine.decode_jpeg(data))


def test_jpeg_encoder_stream_is_pinned_and_generic_equals_avx2_form(tmp_path):
    """pe_encode_jpeg's byte stream on fixed images (the hashes are those of the round's first, scalar encoder: the faster DCT /
    bit writer changed no byte), and the generic instance of the DCT + quantiser (a build without the AVX2 one) writes the same bytes
    as the library on random images of every small size and quality."""
    import ctypes as C
    import hashlib

    from caffe_rtpose_b200 import synth
    pinned = {(96, 160, 98): "99d5b9ec7fc4373bd00a45c6cc10501666c0286699dcb912873c27424ffe6012",
              (33, 70, 50): "bae0e1ac675574e862c7ad234608388c1c7c073f683351e5ece63a6a64429a13"}
    for (h, w, q), want in pinned.items():
        assert hashlib.sha256(engine.encode_jpeg(synth.make_frame(3, h, w), q)).hexdigest() == want
    so = str(tmp_path / "libjpegenc_generic.so")
    src = os.path.join(ROOT, "caffe_rtpose_b200", "csrc", "jpeg.cpp")
    r = subprocess.run(["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-DPE_JPEG_NO_AVX2", "-I", os.path.join(ROOT, "include"), src, "-o", so],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    alt = C.CDLL(so)
    alt.pe_encode_jpeg.restype = C.c_longlong
    alt.pe_encode_jpeg.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong]
    rng = np.random.default_rng(17)
    for i in range(120):
        h, w, q = int(rng.integers(1, 60)), int(rng.integers(1, 80)), int(rng.integers(1, 101))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if i % 3 else np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8)
        buf = np.zeros(h * w * 3 + 70000, np.uint8)
        n = alt.pe_encode_jpeg(img.ctypes.data, w, h, q, buf.ctypes.data, buf.size)
        assert n > 0 and buf[:n].tobytes() == engine.encode_jpeg(img, q), (h, w, q)


def _png_chunk(tag, body):
    import struct
    import zlib
    return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)


def _png_filter_row(kind, cur, prev, bpp):
    cur = np.frombuffer(cur, np.uint8).astype(np.int32)
    prev = np.frombuffer(prev, np.uint8).astype(np.int32)
    left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]]) if len(cur) > bpp else np.zeros_like(cur)
    upleft = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]]) if len(cur) > bpp else np.zeros_like(cur)
    if kind == 0:
        pred = 0
    elif kind == 1:
        pred = left
    elif kind == 2:
        pred = prev
    elif kind == 3:
        pred = (left + prev) // 2
    else:
        p = left + prev - upleft
        pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - upleft)
        pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
    return bytes([kind]) + ((cur - pred) & 255).astype(np.uint8).tobytes()


def _make_png(samples, depth, ctype, palette=None, interlace=False, idat_split=1, trns=None, level=6):
    """samples: (H, W, channels) integer array of raw sample values at `depth` bits; filters cycle 0..4 over the rows."""
    import struct
    import zlib
    H, W, ch = samples.shape

    def pack_rows(sub):
        h, w, _ = sub.shape
        out, prev = b"", None
        bits = ch * depth
        bpp = max(bits // 8, 1)
        for y in range(h):
            v = sub[y].reshape(-1)
            if depth == 16:
                row = v.astype(">u2").tobytes()
            elif depth == 8:
                row = v.astype(np.uint8).tobytes()
            else:
                per = 8 // depth
                pad = (-len(v)) % per
                vv = np.concatenate([v, np.zeros(pad, v.dtype)]).reshape(-1, per)
                row = bytes(int(sum(int(s) << ((per - 1 - i) * depth) for i, s in enumerate(r))) for r in vv)
            if prev is None:
                prev = bytes(len(row))
            out += _png_filter_row(y % 5, row, prev, bpp)
            prev = row
        return out

    if not interlace:
        raw = pack_rows(samples)
    else:
        raw = b""
        for xs, ys, dx, dy in [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]:
            sub = samples[ys::dy, xs::dx]
            if sub.shape[0] and sub.shape[1]:
                raw += pack_rows(sub)
    z = zlib.compress(raw, level)
    n = max(len(z) // idat_split, 1)
    out = b"\x89PNG\r\n\x1a\n" + _png_chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, depth, ctype, 0, 0, 1 if interlace else 0))
    if palette is not None:
        out += _png_chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    if trns is not None:
        out += _png_chunk(b"tRNS", bytes(trns))
    for i in range(0, len(z), n):
        out += _png_chunk(b"IDAT", z[i:i + n])
    return out + _png_chunk(b"IEND", b"")


def test_png_decoder_equals_cv_imread():
    """pe_decode_png feeds --image_dir where the reference calls cv::imread on .png files (rtpose.cpp:1743, :302-391): the
    same 8-bit BGR pixels as OpenCV's libpng path, for the files cv2 can write (8/16-bit BGR, BGRA, grey, bilevel, every zlib
    strategy/level) and for hand-built palette, low-bit-depth, grey+alpha, Adam7-interlaced, all-filter, split-IDAT files."""
    import cv2
    from caffe_rtpose_b200 import synth

    def same(data, what):
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        assert ref is not None, what
        assert np.array_equal(engine.decode_png(data), ref), what

    for h, w in [(72, 128), (17, 33), (1, 1), (5, 3), (9, 8)]:
        img = synth.make_frame(3, h, w)
        img = cv2.GaussianBlur(img, (0, 0), 2) if min(h, w) > 8 else img
        for lvl in (0, 1, 9):
            same(cv2.imencode(".png", img, [cv2.IMWRITE_PNG_COMPRESSION, lvl])[1].tobytes(), ("bgr8", h, w, lvl))
        for strat in (cv2.IMWRITE_PNG_STRATEGY_FILTERED, cv2.IMWRITE_PNG_STRATEGY_HUFFMAN_ONLY, cv2.IMWRITE_PNG_STRATEGY_RLE, cv2.IMWRITE_PNG_STRATEGY_FIXED):
            same(cv2.imencode(".png", img, [cv2.IMWRITE_PNG_STRATEGY, strat])[1].tobytes(), ("strategy", strat))
        same(cv2.imencode(".png", img[:, :, 0])[1].tobytes(), "grey8")
        same(cv2.imencode(".png", np.dstack([img, img[:, :, 0]]))[1].tobytes(), "bgra8")
        im16 = img.astype(np.uint16) * 257 + (img[:, :, ::-1].astype(np.uint16) % 200)
        same(cv2.imencode(".png", im16)[1].tobytes(), "bgr16")
        same(cv2.imencode(".png", im16[:, :, 0])[1].tobytes(), "grey16")
        same(cv2.imencode(".png", (img[:, :, 0] > 128).astype(np.uint8) * 255, [cv2.IMWRITE_PNG_BILEVEL, 1])[1].tobytes(), "bilevel")
    rng = np.random.default_rng(5)
    for h, w in [(13, 21), (8, 8), (1, 7), (20, 3)]:
        for interlace in (False, True):
            for depth in (1, 2, 4, 8):
                pal = rng.integers(0, 256, (1 << depth, 3))
                idx = rng.integers(0, 1 << depth, (h, w, 1))
                same(_make_png(idx, depth, 3, palette=pal, interlace=interlace), ("palette", depth, interlace, h, w))
                same(_make_png(idx, depth, 3, palette=pal, interlace=interlace, trns=[0, 128]), ("palette+tRNS", depth))
                same(_make_png(idx, depth, 0, interlace=interlace), ("grey", depth, interlace))
            for depth in (8, 16):
                top = 1 << depth
                same(_make_png(rng.integers(0, top, (h, w, 3)), depth, 2, interlace=interlace, idat_split=3), ("rgb", depth, interlace))
                same(_make_png(rng.integers(0, top, (h, w, 4)), depth, 6, interlace=interlace), ("rgba", depth, interlace))
                same(_make_png(rng.integers(0, top, (h, w, 2)), depth, 4, interlace=interlace, level=0), ("grey+alpha", depth, interlace))
                same(_make_png(rng.integers(0, top, (h, w, 1)), depth, 0, interlace=interlace), ("grey", depth, interlace))
    with pytest.raises(engine.PoseEngineError):
        engine.decode_png(b"\x89PNG\r\n\x1a\n" + bytes(40))
    good = cv2.imencode(".png", synth.make_frame(1, 16, 16))[1].tobytes()
    with pytest.raises(engine.PoseEngineError):
        engine.decode_png(good[:len(good) // 2])


def test_parsers_survive_corrupt_files(tmp_path):
    """The parsers that read untrusted files (JPEG, PNG, .caffemodel, AVI) under AddressSanitizer + UBSan on thousands of
    mutated files: they must accept or reject, never read out of bounds / overflow / crash."""
    import cv2
    from caffe_rtpose_b200 import synth
    img = cv2.GaussianBlur(synth.make_frame(2, 61, 83), (0, 0), 1.5)
    files = []
    for name, params in [("a.jpg", [cv2.IMWRITE_JPEG_QUALITY, 80]), ("b.jpg", [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_PROGRESSIVE, 1]),
                         ("c.jpg", [cv2.IMWRITE_JPEG_QUALITY, 60, cv2.IMWRITE_JPEG_RST_INTERVAL, 2]), ("d.png", []),
                         ("e.png", [cv2.IMWRITE_PNG_BILEVEL, 1])]:
        data = cv2.imencode("." + name.split(".")[1], img[:, :, 0] if name == "e.png" else img, params)[1].tobytes()
        (tmp_path / name).write_bytes(data)
        files.append(str(tmp_path / name))
    rng = np.random.default_rng(1)
    W = {"conv1_1": (rng.standard_normal((4, 3, 3, 3)).astype(np.float32), np.zeros(4, np.float32)),
         "conv1_2": (rng.standard_normal((5, 4, 1, 1)).astype(np.float32), np.ones(5, np.float32))}
    engine.write_caffemodel(str(tmp_path / "m.caffemodel"), W, [("conv1_1", 4, 3, 3), ("conv1_2", 5, 4, 1)])
    files.append(str(tmp_path / "m.caffemodel"))
    # AVI reader: a Motion-JPEG file and an uncompressed one (tiny frames keep the mutated headers' claims decodable)
    small = cv2.imencode(".jpg", img[:24, :32], [cv2.IMWRITE_JPEG_QUALITY, 75])[1].tobytes()
    _write_avi(str(tmp_path / "v.avi"), [(small, (32, 24))] * 3, b"MJPG")
    dib = b"".join(bytes(img[y, :9].tobytes()).ljust(28, b"\0") for y in range(6, -1, -1))
    _write_avi(str(tmp_path / "w.avi"), [(dib, (9, 7))] * 3, b"DIB ", chunk_tag=b"00db", in_rec=True)
    files += [str(tmp_path / "v.avi"), str(tmp_path / "w.avi")]
    src = os.path.join(ROOT, "caffe_rtpose_b200", "csrc")
    exe = str(tmp_path / "fuzz_codecs")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                        "-I", os.path.join(ROOT, "include"), "-I", src, "-I", "/usr/local/cuda/include",
                        os.path.join(ROOT, "tests", "fuzz", "fuzz_codecs.cpp"), os.path.join(src, "jpeg_dec.cpp"), os.path.join(src, "png_dec.cpp"),
                        os.path.join(src, "caffemodel.cpp"), os.path.join(src, "video.cpp"), "-o", exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, "1500"] + files, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    acc, rej = [int(v) for v in r.stdout.split()[1::2]]
    assert acc > 500 and rej > 500   # the mutations are neither all harmless nor all fatal


def _riff_chunk(tag, body):
    return tag + struct.pack("<I", len(body)) + body + (b"\0" if len(body) & 1 else b"")


def _write_avi(path, frames, fourcc, fps=(25, 1), bits=24, height_sign=1, chunk_tag=b"00dc", extra_movi=b"", in_rec=False):
    """Minimal AVI 1.0 writer for the tests (the RIFF structure AVI files share): hdrl(avih, strl(strh, strf)) + movi."""
    import struct as st
    w, h = frames and frames[0][1] or (16, 16)
    payloads = [f[0] for f in frames]
    avih = st.pack("<IIIIIIIIIIIIII", 1000000 * fps[1] // fps[0], 0, 0, 0x10, len(payloads), 0, 1, 0, w, h, 0, 0, 0, 0)
    strh = b"vids" + fourcc + st.pack("<IHHIIIIIIIIHHHH", 0, 0, 0, 0, fps[1], fps[0], 0, len(payloads), 0, 0xFFFFFFFF, 0, 0, 0, w, h)
    strf = st.pack("<IiiHH4sIiiII", 40, w, h * height_sign, 1, bits, fourcc if fourcc not in (b"DIB ", b"\0\0\0\0") else b"\0\0\0\0", 0, 0, 0, 0, 0)
    aud_strh = b"auds" + b"\0" * 44          # a second (audio) stream must not disturb the frame index
    hdrl = b"hdrl" + _riff_chunk(b"avih", avih) + _riff_chunk(b"LIST", b"strl" + _riff_chunk(b"strh", strh) + _riff_chunk(b"strf", strf)) + \
        _riff_chunk(b"LIST", b"strl" + _riff_chunk(b"strh", aud_strh) + _riff_chunk(b"strf", b"\0" * 16))
    chunks = b"".join(_riff_chunk(chunk_tag, p) + _riff_chunk(b"01wb", b"\1\2\3") for p in payloads)
    if in_rec:
        chunks = _riff_chunk(b"LIST", b"rec " + chunks)
    movi = b"movi" + chunks + extra_movi
    body = b"AVI " + _riff_chunk(b"LIST", hdrl) + _riff_chunk(b"JUNK", b"\0" * 13) + _riff_chunk(b"LIST", movi) + _riff_chunk(b"idx1", b"\0" * 16)
    with open(path, "wb") as f:
        f.write(b"RIFF" + st.pack("<I", len(body)) + body)


def test_video_reader_equals_opencv(tmp_path):
    """--video (cv::VideoCapture, rtpose.cpp:394-411): AVI files written by OpenCV itself (its FFmpeg and its built-in MJPEG writer)
    read back frame for frame - count, size, fps as cv2.VideoCapture reports them, pixels bit-identical to OpenCV's own MJPEG reader
    (imdecode = libjpeg arithmetic) - plus seeking, the end of the file and the error paths."""
    import cv2
    frames = [synth.make_frame(i, 120, 176) for i in range(6)]
    for api in (cv2.CAP_OPENCV_MJPEG, cv2.CAP_FFMPEG):
        path = str(tmp_path / ("v%d.avi" % api))
        wr = cv2.VideoWriter(path, api, cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (176, 120))
        if not wr.isOpened():
            continue
        for f in frames:
            wr.write(f)
        wr.release()
        ref = cv2.VideoCapture(path, cv2.CAP_OPENCV_MJPEG)
        cap = engine.VideoCapture(path)
        assert cap.isOpened(), cap.error
        assert cap.fourcc == "MJPG" and (cap.width, cap.height) == (176, 120)
        assert cap.get(cap.CAP_PROP_FRAME_COUNT) == ref.get(cv2.CAP_PROP_FRAME_COUNT) == len(frames)
        assert abs(cap.get(cap.CAP_PROP_FPS) - ref.get(cv2.CAP_PROP_FPS)) < 1e-9
        for i in range(len(frames)):
            ok, got = cap.read()
            rok, want = ref.read()
            assert ok and rok and np.array_equal(got, want), (api, i)
            assert int(np.argmin([np.abs(got.astype(int) - f.astype(int)).mean() for f in frames])) == i   # and it is the frame that was written
        assert cap.read() == (False, None) and cap.get(cap.CAP_PROP_POS_FRAMES) == len(frames)
        cap.set(cap.CAP_PROP_POS_FRAMES, 3)                                       # FLAGS_start_frame / looping (rtpose.cpp:409-411, 541-543)
        ref.set(cv2.CAP_PROP_POS_FRAMES, 3)
        assert np.array_equal(cap.read()[1], ref.read()[1])
        cap.release()
    # error paths: missing file, not an AVI, an inter-frame codec
    cap = engine.VideoCapture(str(tmp_path / "missing.avi"))
    assert not cap.isOpened() and "Couldn't open video file" in cap.error
    (tmp_path / "x.mp4").write_bytes(b"\0\0\0\x18ftypmp42" + b"\0" * 64)
    cap = engine.VideoCapture(str(tmp_path / "x.mp4"))
    assert not cap.isOpened() and "not a RIFF AVI" in cap.error
    _write_avi(str(tmp_path / "h264.avi"), [(b"\0\0\0\1abc", (176, 120))], b"H264")
    cap = engine.VideoCapture(str(tmp_path / "h264.avi"))
    assert not cap.isOpened() and "H264" in cap.error


def test_video_reader_avi_structure_variants(tmp_path):
    """Hand-built AVI files: uncompressed DIB frames (bottom-up and top-down, 24 and 32 bit, rows padded to 4 bytes), 'rec ' lists,
    a second stream, JUNK chunks, odd-sized chunks, dropped (empty) frames, Motion-JPEG frames without DHT, a truncated file."""
    import cv2
    rng = np.random.default_rng(5)
    w, h = 13, 7                                   # 39 bytes per row -> padded to 40
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(4)]

    def dib(img, bits, top_down):
        px = img if bits == 24 else np.concatenate([img, np.full((h, w, 1), 255, np.uint8)], axis=2)
        stride = (w * bits // 8 + 3) & ~3
        rows = [px[y].tobytes().ljust(stride, b"\0") for y in (range(h) if top_down else range(h - 1, -1, -1))]
        return b"".join(rows)

    for bits in (24, 32):
        for top_down in (False, True):
            for in_rec in (False, True):
                path = str(tmp_path / ("dib%d_%d_%d.avi" % (bits, top_down, in_rec)))
                payloads = [(dib(im, bits, top_down), (w, h)) for im in imgs]
                payloads.insert(2, (b"", (w, h)))   # dropped frame: repeats frame 1
                _write_avi(path, payloads, b"DIB ", fps=(30000, 1001), bits=bits, height_sign=-1 if top_down else 1, chunk_tag=b"00db", in_rec=in_rec)
                cap = engine.VideoCapture(path)
                assert cap.isOpened(), cap.error
                assert (cap.width, cap.height, cap.frame_count, cap.fourcc) == (w, h, 5, "DIB ") and abs(cap.fps - 30000 / 1001) < 1e-9
                want = [imgs[0], imgs[1], imgs[1], imgs[2], imgs[3]]
                for i in range(5):
                    ok, got = cap.read()
                    assert ok and np.array_equal(got, want[i]), (bits, top_down, in_rec, i)
    # the uncompressed file OpenCV's FFmpeg writer produces (fourcc 0) reads back identically
    path = str(tmp_path / "raw.avi")
    wr = cv2.VideoWriter(path, cv2.CAP_FFMPEG, 0, 10.0, (w + 3, h + 1))
    if wr.isOpened():
        big = [rng.integers(0, 256, (h + 1, w + 3, 3), dtype=np.uint8) for _ in range(3)]
        for f in big:
            wr.write(f)
        wr.release()
        cap = engine.VideoCapture(path)
        if cap.isOpened():                           # FFmpeg may pick a pixel format other than bgr24; then the reader must refuse
            for f in big:
                assert np.array_equal(cap.read()[1], f)
    # Motion-JPEG frames without Huffman tables (what cameras and many capture tools write): the Annex K tables apply
    img = synth.make_frame(3, 64, 96)
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])
    data = enc.tobytes()
    out, i = bytearray(data[:2]), 2
    while i < len(data):                             # drop the DHT segments; quality-90 baseline files use the standard tables
        assert data[i] == 0xFF
        m, ln = data[i + 1], struct.unpack(">H", data[i + 2:i + 4])[0]
        if m == 0xDA:
            out += data[i:]
            break
        if m != 0xC4:
            out += data[i:i + 2 + ln]
        i += 2 + ln
    assert len(out) < len(data) - 400
    assert np.array_equal(engine.decode_jpeg(bytes(out)), cv2.imdecode(enc, cv2.IMREAD_COLOR))
    path = str(tmp_path / "nodht.avi")
    _write_avi(path, [(bytes(out), (96, 64))] * 3, b"MJPG")
    cap = engine.VideoCapture(path)
    assert cap.isOpened() and np.array_equal(cap.read()[1], cv2.imdecode(enc, cv2.IMREAD_COLOR))
    # OpenDML: frames continue in 'AVIX' extension RIFFs after the first one (files > 1 GB); foreign RIFFs in between are skipped
    import struct as st
    first = open(path, "rb").read()
    more = b"AVIX" + _riff_chunk(b"LIST", b"movi" + _riff_chunk(b"00dc", bytes(out)) + _riff_chunk(b"01wb", b"\0\0") + _riff_chunk(b"00dc", bytes(out)))
    odml = str(tmp_path / "odml.avi")
    open(odml, "wb").write(first + b"RIFF" + st.pack("<I", 12) + b"WAVEfmt \0\0\0\0" + b"RIFF" + st.pack("<I", len(more)) + more)
    cap = engine.VideoCapture(odml)
    assert cap.isOpened() and cap.frame_count == 5
    cap.set(cap.CAP_PROP_POS_FRAMES, 4)
    assert np.array_equal(cap.read()[1], cv2.imdecode(enc, cv2.IMREAD_COLOR))
    # a recording that was cut off: the complete frames stay readable
    whole = open(path, "rb").read()
    cut = str(tmp_path / "cut.avi")
    open(cut, "wb").write(whole[:len(whole) - len(out) - 40])
    cap = engine.VideoCapture(cut)
    assert cap.isOpened() and cap.frame_count == 2 and cap.read()[0] and cap.read()[0] and not cap.read()[0]


def test_camera_conversion_equals_opencv_and_open_fails_loudly():
    """Camera frames (rtpose.cpp:401-405, 431): YUYV -> BGR with cv::cvtColor(COLOR_YUV2BGR_YUYV)'s arithmetic (what OpenCV's V4L2
    back end applies), bit for bit incl. the saturating corners; opening a device that does not exist reports the reference's message."""
    import cv2
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (36, 50, 2), dtype=np.uint8)
    img[0, :8] = [[0, 0], [255, 255], [16, 128], [235, 128], [0, 255], [255, 0], [81, 90], [81, 240]]
    assert np.array_equal(engine.yuyv_to_bgr(img), cv2.cvtColor(img, cv2.COLOR_YUV2BGR_YUYV))
    grid = np.stack(np.meshgrid(np.arange(0, 256, 5), np.arange(0, 256, 3), indexing="ij"), -1).astype(np.uint8)   # (Y, U|V) sweep
    grid = grid[:, :grid.shape[1] // 2 * 2]
    assert np.array_equal(engine.yuyv_to_bgr(grid), cv2.cvtColor(np.ascontiguousarray(grid), cv2.COLOR_YUV2BGR_YUYV))
    with pytest.raises(engine.PoseEngineError):
        engine.yuyv_to_bgr(np.zeros((4, 5, 2), np.uint8))          # odd width
    h = C.c_void_p()
    assert engine.lib().pe_camera_open(63, 1280, 720, C.byref(h)) != 0 and not h.value
    assert b"Couldn't open camera 63 (/dev/video63" in engine.lib().pe_camera_last_error()
