// Model descriptors and the engine's execution plan for the deploy graph.
//
// Reference: src/rtpose/modelDescriptorFactory.cpp:6-28 (MPI_15), 30-55 (COCO_18);
//            model/{coco,mpi}/pose_deploy_linevec.prototxt (92 conv, 80 ReLU, 3 pool, 6 concat).
// The plan is not a translation of Caffe's layer list: ReLU is fused into the conv epilogue, Concat is
// removed (producers write channel slices of a shared buffer), Split disappears, and conv1_1 consumes an
// im2col'ed input so that it is a K=27 1x1 GEMM.
#include <stdio.h>

#include "common.h"

namespace pe {

static const int kLimbMPI[] = {0, 1, 1, 2, 2, 3, 3, 4, 1, 5, 5, 6, 6, 7, 1, 14, 14, 11, 11, 12, 12, 13, 14, 8, 8, 9, 9, 10};
static const int kMapMPI[] = {16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 38, 39, 40, 41, 42, 43, 32, 33, 34, 35, 36, 37};
static const int kLimbCOCO[] = {1, 2, 1, 5, 2, 3, 3, 4, 5, 6, 6, 7, 1, 8, 8, 9, 9, 10, 1, 11, 11, 12, 12, 13, 1, 0, 0, 14, 14, 16, 0, 15, 15, 17, 2, 16, 5, 17};
static const int kMapCOCO[] = {31, 32, 39, 40, 33, 34, 35, 36, 41, 42, 43, 44, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 47, 48, 49, 50, 53, 54, 51, 52, 55, 56, 37, 38, 45, 46};
static const char* kNamesMPI[] = {"Head", "Neck", "RShoulder", "RElbow", "RWrist", "LShoulder", "LElbow", "LWrist", "RHip",
                                  "RKnee", "RAnkle", "LHip", "LKnee", "LAnkle", "Chest", "Bkg"};
static const char* kNamesCOCO[] = {"Nose", "Neck", "RShoulder", "RElbow", "RWrist", "LShoulder", "LElbow", "LWrist", "RHip", "RKnee",
                                   "RAnkle", "LHip", "LKnee", "LAnkle", "REye", "LEye", "REar", "LEar", "Bkg"};

const ModelTables& model_tables(int model) {
    // max_peaks / num_parts: nms_param of the deploy prototxts (coco :2989-3000, mpi :2976-2999)
    static const ModelTables mpi = {15, 14, 44, 20, kLimbMPI, kMapMPI};
    static const ModelTables coco = {18, 19, 57, 64, kLimbCOCO, kMapCOCO};
    return model == PE_MODEL_MPI_15 ? mpi : coco;
}

// modelDescriptor.cpp:4-20 createPartToName: PAF channels are named "<A>-><B>(X|Y)"
const char* model_part_name(int model, int idx) {
    static std::vector<std::string> names[2];
    const int mi = model == PE_MODEL_MPI_15 ? 0 : 1;
    const ModelTables& t = model_tables(model);
    if (names[mi].empty()) {
        names[mi].assign(t.num_maps, "");
        for (int i = 0; i <= t.num_parts; i++) names[mi][i] = mi == 0 ? kNamesMPI[i] : kNamesCOCO[i];
        for (int l = 0; l < t.num_limbs; l++) {
            const std::string base = names[mi][t.limb_seq[2 * l]] + "->" + names[mi][t.limb_seq[2 * l + 1]];
            names[mi][t.map_idx[2 * l]] = base + "(X)";
            names[mi][t.map_idx[2 * l + 1]] = base + "(Y)";
        }
    }
    if (idx < 0 || idx >= t.num_maps) return "";
    return names[mi][idx].c_str();
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

NetPlan build_plan(int model, int kp_input, int cpad) {
    NetPlan p;
    const ModelTables& mt = model_tables(model);
    p.model = model;
    p.c_l1 = 2 * mt.num_limbs;       // PAF branch (L1)
    p.c_l2 = mt.num_parts + 1;       // part + background branch (L2)
    p.kp_input = kp_input;

    auto new_act = [&](int level, int c, const std::string& blob, int blob_c) {
        ActSpec a; a.level = level; a.C = round_up(c, cpad); a.blob = blob; a.blob_c = blob_c;
        p.acts.push_back(a);
        if (!blob.empty()) p.blobs.push_back({blob, (int)p.acts.size() - 1, 0, blob_c});
        return (int)p.acts.size() - 1;
    };
    auto add_conv = [&](const std::string& name, int in_act, int in_cused, std::vector<int> cin_map, int cin, int cout,
                        int k, int relu, int level, int out_act, int out_coff, int planar_coff, int im2col) {
        ConvSpec c;
        c.name = name; c.cout = cout; c.cin = cin; c.k = k; c.pad = k / 2; c.relu = relu; c.level = level;
        c.in_act = in_act; c.in_cused = in_cused; c.out_act = out_act; c.out_coff = out_coff;
        c.planar_coff = planar_coff; c.cin_map = cin_map; c.im2col_input = im2col;
        c.flops_per_image = 0;
        p.convs.push_back(c);
        p.order.push_back({0, (int)p.convs.size() - 1});
        return (int)p.convs.size() - 1;
    };
    auto ident = [](int n, int padded) {
        std::vector<int> m(padded, -1);
        for (int i = 0; i < n; i++) m[i] = i;
        return m;
    };

    // network input, im2col'ed 3x3x3 patches: engine channel (r*3+s)*3+c  <-  original weight index (c, r, s)
    p.input_act = new_act(0, kp_input, "", 0);
    p.acts[p.input_act].C = kp_input;
    p.blobs.push_back({"image", p.input_act, 12, 3});  // centre tap (r=1,s=1) of the patch = the net input itself
    int cur = p.input_act, cur_c = 3, level = 0;
    char nm[64];
    const int vgg[4][2] = {{64, 2}, {128, 2}, {256, 4}, {512, 2}};
    for (int b = 0; b < 4; b++) {
        for (int i = 1; i <= vgg[b][1]; i++) {
            snprintf(nm, sizeof nm, "conv%d_%d", b + 1, i);
            const int out = new_act(level, vgg[b][0], nm, vgg[b][0]);
            if (b == 0 && i == 1) {
                add_conv(nm, cur, kp_input, std::vector<int>(), 3, 64, 3, 1, level, out, 0, 0, 1);
            } else {
                add_conv(nm, cur, p.acts[cur].C, ident(cur_c, p.acts[cur].C), cur_c, vgg[b][0], 3, 1, level, out, 0, 0, 0);
            }
            cur = out; cur_c = vgg[b][0];
        }
        if (b < 3) {
            snprintf(nm, sizeof nm, "pool%d_stage1", b + 1);
            const int out = new_act(level + 1, cur_c, nm, cur_c);
            p.pools.push_back({nm, cur, out, level});
            p.order.push_back({1, (int)p.pools.size() - 1});
            cur = out; level++;
        }
    }
    {
        const int a43 = new_act(3, 256, "conv4_3_CPM", 256);
        add_conv("conv4_3_CPM", cur, p.acts[cur].C, ident(512, p.acts[cur].C), 512, 256, 3, 1, 3, a43, 0, 0, 0);
        cur = a43;
    }
    // two concat buffers, ping-ponged by stage: channels [F 128 | L1 slot | L2 slot | pad]
    const int slot1 = round_up(p.c_l1, 8), slot2 = round_up(p.c_l2, 8);
    const int off_l1 = 128, off_l2 = 128 + slot1, cc_c = round_up(128 + slot1 + slot2, 64);
    int cc[2];
    cc[0] = new_act(3, cc_c, "conv4_4_CPM", 128);
    cc[1] = new_act(3, cc_c, "", 0);
    p.acts[cc[0]].C = p.acts[cc[1]].C = cc_c;
    add_conv("conv4_4_CPM", cur, p.acts[cur].C, ident(256, p.acts[cur].C), 256, 128, 3, 1, 3, cc[0], 0, 0, 0);
    p.copies.push_back({cc[0], cc[1], 128});
    p.order.push_back({2, 0});

    // Caffe concat order of concat_stage{2..6}: [L1 | L2 | F]  (prototxt :731-741)
    std::vector<int> cc_map(cc_c, -1);
    for (int i = 0; i < 128; i++) cc_map[i] = p.c_l1 + p.c_l2 + i;
    for (int i = 0; i < p.c_l1; i++) cc_map[off_l1 + i] = i;
    for (int i = 0; i < p.c_l2; i++) cc_map[off_l2 + i] = p.c_l1 + i;

    // stage 1: reads F = first 128 channels of cc[0]; its last layers write the L1/L2 slices of cc[0]
    for (int br = 1; br <= 2; br++) {
        int in = cc[0], in_cused = 128, in_c = 128;
        std::vector<int> in_map = ident(128, 128);
        for (int i = 1; i <= 5; i++) {
            snprintf(nm, sizeof nm, "conv5_%d_CPM_L%d", i, br);
            if (i <= 3) {
                const int out = new_act(3, 128, nm, 128);
                add_conv(nm, in, in_cused, in_map, in_c, 128, 3, 1, 3, out, 0, 0, 0);
                in = out; in_cused = p.acts[out].C; in_c = 128; in_map = ident(128, in_cused);
            } else if (i == 4) {
                const int out = new_act(3, 512, nm, 512);
                add_conv(nm, in, in_cused, in_map, in_c, 512, 1, 1, 3, out, 0, 0, 0);
                in = out; in_cused = p.acts[out].C; in_c = 512; in_map = ident(512, in_cused);
            } else {
                const int co = br == 1 ? p.c_l1 : p.c_l2;
                add_conv(nm, in, in_cused, in_map, in_c, co, 1, 0, 3, cc[0], br == 1 ? off_l1 : off_l2, 0, 0);
                p.blobs.push_back({nm, cc[0], br == 1 ? off_l1 : off_l2, co});
            }
        }
    }
    // stages 2..6: read cc[(s)&1], write slices of cc[(s+1)&1]; stage 6 writes the final planar maps,
    // concat_stage7 = [L2 | L1]  (prototxt :2966-2975)
    for (int s = 2; s <= 6; s++) {
        const int src = cc[s & 1], dst = cc[(s + 1) & 1];
        for (int br = 1; br <= 2; br++) {
            int in = src, in_cused = cc_c, in_c = p.c_l1 + p.c_l2 + 128;
            std::vector<int> in_map = cc_map;
            for (int i = 1; i <= 7; i++) {
                snprintf(nm, sizeof nm, "Mconv%d_stage%d_L%d", i, s, br);
                if (i <= 6) {
                    const int out = new_act(3, 128, nm, 128);
                    add_conv(nm, in, in_cused, in_map, in_c, 128, i <= 5 ? 7 : 1, 1, 3, out, 0, 0, 0);
                    in = out; in_cused = p.acts[out].C; in_c = 128; in_map = ident(128, in_cused);
                } else {
                    const int co = br == 1 ? p.c_l1 : p.c_l2;
                    if (s < 6) {
                        add_conv(nm, in, in_cused, in_map, in_c, co, 1, 0, 3, dst, br == 1 ? off_l1 : off_l2, 0, 0);
                        p.blobs.push_back({nm, dst, br == 1 ? off_l1 : off_l2, co});
                    } else {
                        add_conv(nm, in, in_cused, in_map, in_c, co, 1, 0, 3, -1, 0, br == 1 ? p.c_l2 : 0, 0);
                    }
                }
            }
        }
    }
    return p;
}

}  // namespace pe
