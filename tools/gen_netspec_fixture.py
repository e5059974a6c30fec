#!/usr/bin/env python3
"""Parse the reference deploy prototxts into compact JSON layer tables.

Run in the build container only (reads /root/reference, which does not exist on
the GPU box).  Output: tests/golden/netspec_{coco,mpi,mpi_1,mpi_2,mpi_4}.json, which the CPU test
suite compares with the oracle's and the engine's built-in graph builders
(reference: model/{coco,mpi}/pose_deploy_linevec.prototxt).

The parser handles the protobuf text-format subset those files use:
`key: value`, `key { ... }`, `#` comments, repeated keys.
"""
import json
import os
import re
import sys

REF = "/root/reference/model"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def tokenize(text):
    text = re.sub(r"#[^\n]*", "", text)
    return re.findall(r'"[^"]*"|[{}]|[^\s{}:]+|:', text)


def parse_block(tok, i):
    """Return (list of (key, value) pairs, next index); value is str or nested list."""
    out = []
    while i < len(tok) and tok[i] != "}":
        key = tok[i]
        i += 1
        if tok[i] == ":":
            i += 1
            if tok[i] == "{":
                val, i = parse_block(tok, i + 1)
                i += 1
            else:
                val = tok[i].strip('"')
                i += 1
        elif tok[i] == "{":
            val, i = parse_block(tok, i + 1)
            i += 1
        else:
            raise ValueError("bad token after key %r: %r" % (key, tok[i]))
        out.append((key, val))
    return out, i


def get(block, key, default=None):
    for k, v in block:
        if k == key:
            return v
    return default


def getall(block, key):
    return [v for k, v in block if k == key]


def to_spec(path):
    top, _ = parse_block(tokenize(open(path).read()), 0)
    spec = {"input": get(top, "input"), "input_dim": [int(v) for v in getall(top, "input_dim")], "layers": []}
    for lay in getall(top, "layer"):
        d = {"name": get(lay, "name"), "type": get(lay, "type"),
             "bottom": getall(lay, "bottom"), "top": getall(lay, "top")}
        cp = get(lay, "convolution_param")
        if cp is not None:
            d.update(num_output=int(get(cp, "num_output")), pad=int(get(cp, "pad", 0)),
                     kernel_size=int(get(cp, "kernel_size")), stride=int(get(cp, "stride", 1)),
                     weight_filler=dict(get(cp, "weight_filler")), bias_filler=dict(get(cp, "bias_filler")))
        pp = get(lay, "pooling_param")
        if pp is not None:
            d.update(pool=get(pp, "pool"), kernel_size=int(get(pp, "kernel_size")),
                     stride=int(get(pp, "stride", 1)), pad=int(get(pp, "pad", 0)))
        ccp = get(lay, "concat_param")
        if ccp is not None:
            d.update(axis=int(get(ccp, "axis", 1)))
        ip = get(lay, "imresize_param")
        if ip is not None:
            d.update(factor=float(get(ip, "factor")), scale_gap=float(get(ip, "scale_gap")),
                     start_scale=float(get(ip, "start_scale")))
        npar = get(lay, "nms_param")
        if npar is not None:
            # NmsParameter defaults (caffe.proto:1471-1476): max_peaks 20, num_parts 15
            d.update(threshold=float(get(npar, "threshold", 0.5)), max_peaks=int(get(npar, "max_peaks", 20)),
                     num_parts=int(get(npar, "num_parts", 15)), max_peaks_default=get(npar, "max_peaks") is None)
        spec["layers"].append(d)
    return spec


def main():
    os.makedirs(OUT, exist_ok=True)
    for model, fname, out in (("coco", "pose_deploy_linevec", "coco"), ("mpi", "pose_deploy_linevec", "mpi"),
                              ("mpi", "pose_deploy_linevec_1", "mpi_1"), ("mpi", "pose_deploy_linevec_2", "mpi_2"),
                              ("mpi", "pose_deploy_linevec_4", "mpi_4")):
        spec = to_spec(os.path.join(REF, model, fname + ".prototxt"))
        model = out
        with open(os.path.join(OUT, "netspec_%s.json" % model), "w") as f:
            json.dump(spec, f, indent=0, sort_keys=True)
        types = {}
        for l in spec["layers"]:
            types[l["type"]] = types.get(l["type"], 0) + 1
        print(model, len(spec["layers"]), types)


if __name__ == "__main__":
    sys.exit(main())
