"""Single-process multi-GPU replica check (the reference's --num_gpu topology, rtpose.cpp:1459-1549): GPU 0 packs the
weights once, pe_broadcast_weights replicates the packed buffer to the other visible GPUs with one grouped ncclBroadcast,
and every replica must then produce bit-identical stride-8 maps, peaks and joints for the same frame.

  python -m caffe_rtpose_b200.replica_check [N]      prints one JSON line, exit code 0 = identical (or fewer than 2 GPUs)

Used by __graft_entry__.smoke() and by rank 0 of `bench.py --gpus N` (in a child process with a timeout, before the
timed region), so that the path is exercised whenever more than one GPU is visible."""
import json
import sys

import numpy as np


def run(n_gpus=None, net_w=160, net_h=96):
    from . import engine, synth
    # visible devices: create engines until the ordinal is refused (pe_create reports "device d out of range")
    model, disp_w, disp_h = engine.COCO_18, 2 * net_w, 2 * net_h
    engs = []
    want = n_gpus or 8
    for d in range(want):
        try:
            engs.append(engine.PoseEngine(model, net_w, net_h, disp_w, disp_h, precision=engine.PREC_F16X2, device=d))
        except engine.PoseEngineError:
            break
    out = {"check": "pe_broadcast_weights", "gpus": len(engs)}
    if len(engs) < 2:
        out["result"] = "skipped: fewer than 2 GPUs visible"
        for e in engs:
            e.close()
        return out, 0
    engs[0].set_weights(synth.make_weights(model, "he"))
    engine.broadcast_weights(engs)                     # the replicas never see the fp32 weights
    frame = synth.make_frame(5, disp_h, disp_w)
    res = []
    for e in engs:
        e.forward_frames([frame])
        cnt, joints, peaks = e.fetch(0)
        res.append((cnt, joints.copy(), peaks.copy(), e.fetch_maps(1).copy()))
    same = all(r[0] == res[0][0] and all(np.array_equal(a, b) for a, b in zip(r[1:], res[0][1:])) for r in res[1:])
    out.update(result="replicas bit-identical" if same else "REPLICAS DIFFER", identical=bool(same), people=int(res[0][0]),
               peaks=int(np.minimum(res[0][2][:, 0, 0], res[0][2].shape[1] - 1).sum()), packed_bytes=int(engs[0].packed_weights()[1]))
    for e in engs:
        e.close()
    return out, 0 if same else 1


if __name__ == "__main__":
    o, rc = run(int(sys.argv[1]) if len(sys.argv) > 1 else None)
    print(json.dumps(o), flush=True)
    sys.exit(rc)
