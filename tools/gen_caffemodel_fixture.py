#!/usr/bin/env python3
"""Independent .caffemodel fixtures for the wire-format reader (csrc/caffemodel.cpp), written by Google's protobuf runtime -
NOT by caffe_rtpose_b200.engine.write_caffemodel - from message descriptors that restate the relevant part of
src/caffe/proto/caffe.proto (field numbers and types as in the reference; protoc is not available here, so the descriptors
are built programmatically):

  NetParameter      name=1, layers=2 (V1LayerParameter), input=3, input_dim=4, force_backward=5, state=6, layer=100
  LayerParameter    name=1, type=2, bottom=3, top=4, loss_weight=5, param=6 (ParamSpec), blobs=7, phase=10, convolution_param=106
  V1LayerParameter  bottom=2, top=3, name=4, type=5 (enum; CONVOLUTION = 4), blobs=6, blobs_lr=7, weight_decay=8
  BlobProto         num=1, channels=2, height=3, width=4 (legacy), data=5 (repeated float, packed), diff=6, shape=7, double_data=8
  BlobShape         dim=1 (repeated int64, packed)

Files (tests/golden/):
  caffemodel_v2.caffemodel        `layer` blocks, BlobShape, packed data, layer fields the reader must skip (bottom/top/param/
                                  convolution_param/phase), a non-conv layer without blobs, a `diff` array
  caffemodel_v1.caffemodel        legacy `layers` blocks (V1 enum type), legacy num/channels/height/width dims, data written
                                  UNPACKED (one tag per float, as old protobuf writers emit), blobs_lr / weight_decay
  caffemodel_fixture.npz          the arrays both files hold
Run in the build container:  python tools/gen_caffemodel_fixture.py"""
import os

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
F = descriptor_pb2.FieldDescriptorProto


def field(msg, name, number, ftype, label=F.LABEL_OPTIONAL, type_name=None, packed=None):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = type_name
    if packed is not None:
        f.options.packed = packed
    return f


def build(pkg, packed_data):
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = pkg + ".proto", pkg, "proto2"
    m = fd.message_type.add(); m.name = "BlobShape"
    field(m, "dim", 1, F.TYPE_INT64, F.LABEL_REPEATED, packed=True)
    m = fd.message_type.add(); m.name = "BlobProto"
    for i, n in enumerate(["num", "channels", "height", "width"]):
        field(m, n, i + 1, F.TYPE_INT32)
    field(m, "data", 5, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=packed_data)
    field(m, "diff", 6, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=packed_data)
    field(m, "shape", 7, F.TYPE_MESSAGE, type_name=".%s.BlobShape" % pkg)
    field(m, "double_data", 8, F.TYPE_DOUBLE, F.LABEL_REPEATED, packed=True)
    m = fd.message_type.add(); m.name = "ParamSpec"
    field(m, "name", 1, F.TYPE_STRING); field(m, "lr_mult", 3, F.TYPE_FLOAT); field(m, "decay_mult", 4, F.TYPE_FLOAT)
    m = fd.message_type.add(); m.name = "ConvolutionParameter"
    field(m, "num_output", 1, F.TYPE_UINT32); field(m, "pad", 3, F.TYPE_UINT32, F.LABEL_REPEATED); field(m, "kernel_size", 4, F.TYPE_UINT32, F.LABEL_REPEATED)
    m = fd.message_type.add(); m.name = "LayerParameter"
    field(m, "name", 1, F.TYPE_STRING); field(m, "type", 2, F.TYPE_STRING)
    field(m, "bottom", 3, F.TYPE_STRING, F.LABEL_REPEATED); field(m, "top", 4, F.TYPE_STRING, F.LABEL_REPEATED)
    field(m, "loss_weight", 5, F.TYPE_FLOAT, F.LABEL_REPEATED)
    field(m, "param", 6, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".%s.ParamSpec" % pkg)
    field(m, "blobs", 7, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".%s.BlobProto" % pkg)
    field(m, "phase", 10, F.TYPE_INT32)
    field(m, "convolution_param", 106, F.TYPE_MESSAGE, type_name=".%s.ConvolutionParameter" % pkg)
    m = fd.message_type.add(); m.name = "V1LayerParameter"
    field(m, "bottom", 2, F.TYPE_STRING, F.LABEL_REPEATED); field(m, "top", 3, F.TYPE_STRING, F.LABEL_REPEATED)
    field(m, "name", 4, F.TYPE_STRING); field(m, "type", 5, F.TYPE_INT32)
    field(m, "blobs", 6, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".%s.BlobProto" % pkg)
    field(m, "blobs_lr", 7, F.TYPE_FLOAT, F.LABEL_REPEATED); field(m, "weight_decay", 8, F.TYPE_FLOAT, F.LABEL_REPEATED)
    m = fd.message_type.add(); m.name = "NetParameter"
    field(m, "name", 1, F.TYPE_STRING)
    field(m, "layers", 2, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".%s.V1LayerParameter" % pkg)
    field(m, "input", 3, F.TYPE_STRING, F.LABEL_REPEATED); field(m, "input_dim", 4, F.TYPE_INT32, F.LABEL_REPEATED)
    field(m, "force_backward", 5, F.TYPE_BOOL)
    field(m, "layer", 100, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".%s.LayerParameter" % pkg)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName(pkg + ".NetParameter"))


def main():
    rng = np.random.default_rng(20260923)
    layers = [("conv1_1", (64, 3, 3, 3)), ("conv5_5_CPM_L2", (19, 512, 1, 1)), ("Mconv7_stage6_L1", (38, 128, 1, 1))]
    arrays = {}
    for name, shp in layers:
        arrays[name + "_w"] = rng.standard_normal(shp).astype(np.float32)
        arrays[name + "_b"] = rng.standard_normal(shp[0]).astype(np.float32)
    # ---- V2: `layer`, BlobShape, packed
    Net = build("caffe_v2", True)
    net = Net()
    net.name = "fixture_v2"
    net.input.append("image"); net.input_dim.extend([1, 3, 368, 656]); net.force_backward = False
    for name, shp in layers:
        l = net.layer.add()
        l.name, l.type = name, "Convolution"
        l.bottom.append("x"); l.top.append(name); l.phase = 1
        for lr, dm in ((1.0, 1.0), (2.0, 0.0)):
            p = l.param.add(); p.lr_mult, p.decay_mult = lr, dm
        l.convolution_param.num_output = shp[0]; l.convolution_param.kernel_size.append(shp[2]); l.convolution_param.pad.append(shp[2] // 2)
        for key, dims in (("_w", shp), ("_b", (shp[0],))):
            b = l.blobs.add()
            b.shape.dim.extend(dims)
            b.data.extend(arrays[name + key].ravel().tolist())
            if key == "_b":
                b.diff.extend([0.0] * shp[0])                 # solver snapshots carry diffs; the reader ignores them
        r = net.layer.add()
        r.name, r.type = "relu_" + name, "ReLU"               # a layer without blobs
        r.bottom.append(name); r.top.append(name)
    open(os.path.join(OUT, "caffemodel_v2.caffemodel"), "wb").write(net.SerializeToString())
    # ---- V1: `layers`, legacy dims, unpacked floats
    Net1 = build("caffe_v1", False)
    net = Net1()
    net.name = "fixture_v1"
    for name, shp in layers:
        l = net.layers.add()
        l.name, l.type = name, 4                               # V1LayerParameter.LayerType CONVOLUTION
        l.bottom.append("x"); l.top.append(name)
        l.blobs_lr.extend([1.0, 2.0]); l.weight_decay.extend([1.0, 0.0])
        for key, dims in (("_w", shp), ("_b", (1, 1, 1, shp[0]))):     # legacy biases are 1 x 1 x 1 x N (blob.cpp:448-470)
            b = l.blobs.add()
            b.num, b.channels, b.height, b.width = dims
            b.data.extend(arrays[name + key].ravel().tolist())
    open(os.path.join(OUT, "caffemodel_v1.caffemodel"), "wb").write(net.SerializeToString())
    np.savez_compressed(os.path.join(OUT, "caffemodel_fixture.npz"), **arrays)
    for f in ("caffemodel_v2.caffemodel", "caffemodel_v1.caffemodel"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
