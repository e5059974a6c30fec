// Protobuf text-format reader for deploy prototxts + the built-in network definitions.  See prototxt.h.
#include "prototxt.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <memory>

#include "common.h"

namespace pe {
namespace {

struct Node;
struct Field {
    std::string key, scalar;          // scalar: unquoted text of `key: value`
    std::unique_ptr<Node> msg;        // `key { ... }`
};
struct Node { std::vector<Field> fields; };

struct Lexer {
    const std::string& s;
    size_t i = 0;
    int line = 1;
    explicit Lexer(const std::string& t) : s(t) {}
    void skip() {
        for (;;) {
            while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\r' || s[i] == '\n' || s[i] == ',' || s[i] == ';')) {
                if (s[i] == '\n') line++;
                i++;
            }
            if (i < s.size() && s[i] == '#') { while (i < s.size() && s[i] != '\n') i++; continue; }
            break;
        }
    }
    // kinds: 0 end, 1 identifier/number, 2 string, '{' '}' ':' '<' '>' '[' ']'
    int next(std::string& tok) {
        skip();
        tok.clear();
        if (i >= s.size()) return 0;
        const char c = s[i];
        if (c == '{' || c == '}' || c == ':' || c == '<' || c == '>' || c == '[' || c == ']') { i++; tok = c; return c; }
        if (c == '"' || c == '\'') {
            i++;
            while (i < s.size() && s[i] != c) {
                if (s[i] == '\\' && i + 1 < s.size()) {
                    i++;
                    const char e = s[i];
                    tok += e == 'n' ? '\n' : e == 't' ? '\t' : e;
                } else {
                    if (s[i] == '\n') line++;
                    tok += s[i];
                }
                i++;
            }
            if (i >= s.size()) return -1;
            i++;
            return 2;
        }
        while (i < s.size() && !strchr(" \t\r\n{}:<>[],;#\"'", s[i])) tok += s[i++];
        return tok.empty() ? -1 : 1;
    }
};

int parse_node(Lexer& lx, Node& n, int closer, int depth, std::string& err) {
    if (depth > 64) { err = "nesting too deep"; return -1; }
    std::string tok;
    for (;;) {
        const size_t save = lx.i;
        int k = lx.next(tok);
        if (k == 0) { if (closer) { err = "unexpected end of file (unbalanced braces)"; return -1; } return 0; }
        if (k == closer) return 0;
        if (k != 1) { err = "line " + std::to_string(lx.line) + ": expected a field name, got '" + tok + "'"; (void)save; return -1; }
        Field f;
        f.key = tok;
        k = lx.next(tok);
        if (k == ':') k = lx.next(tok);
        if (k == '{' || k == '<') {
            f.msg.reset(new Node());
            if (parse_node(lx, *f.msg, k == '{' ? '}' : '>', depth + 1, err)) return -1;
        } else if (k == '[') {   // `key: [a, b]` list of scalars: one field per element
            for (;;) {
                k = lx.next(tok);
                if (k == ']') break;
                if (k != 1 && k != 2) { err = "line " + std::to_string(lx.line) + ": bad list element"; return -1; }
                Field e; e.key = f.key; e.scalar = tok;
                n.fields.push_back(std::move(e));
            }
            continue;
        } else if (k == 1 || k == 2) {
            f.scalar = tok;
        } else {
            err = "line " + std::to_string(lx.line) + ": field '" + f.key + "' has no value";
            return -1;
        }
        n.fields.push_back(std::move(f));
    }
}

const Field* find(const Node& n, const char* key) {
    for (const Field& f : n.fields) if (f.key == key) return &f;
    return nullptr;
}
bool get_int(const Node& n, const char* key, int& v, std::string& err) {
    const Field* f = find(n, key);
    if (!f) return true;
    char* end = nullptr;
    const double d = strtod(f->scalar.c_str(), &end);
    if (f->msg || end == f->scalar.c_str() || *end || d != (double)(long long)d || d < -2147483648.0 || d > 2147483647.0) {
        err = std::string("field '") + key + "': '" + f->scalar + "' is not an integer";
        return false;
    }
    v = (int)d;
    return true;
}
bool get_float(const Node& n, const char* key, float& v, std::string& err) {
    const Field* f = find(n, key);
    if (!f) return true;
    char* end = nullptr;
    const double d = strtod(f->scalar.c_str(), &end);
    if (f->msg || end == f->scalar.c_str() || (*end && strcmp(end, "f"))) { err = std::string("field '") + key + "' is not a number"; return false; }
    v = (float)d;
    return true;
}

// V1LayerParameter.LayerType enum names (caffe.proto V1LayerParameter, upgrade_proto.cpp UpgradeV1LayerType) -> V2 type strings
const char* upgrade_v1_type(const std::string& t) {
    static const char* const map[][2] = {{"CONVOLUTION", "Convolution"}, {"RELU", "ReLU"}, {"POOLING", "Pooling"}, {"CONCAT", "Concat"},
                                         {"SPLIT", "Split"}, {"DROPOUT", "Dropout"}, {"SOFTMAX", "Softmax"}, {"INNER_PRODUCT", "InnerProduct"},
                                         {"LRN", "LRN"}, {"ELTWISE", "Eltwise"}, {"DECONVOLUTION", "Deconvolution"}, {"SIGMOID", "Sigmoid"},
                                         {"TANH", "TanH"}, {"DATA", "Data"}, {"SLICE", "Slice"}, {"FLATTEN", "Flatten"}, {"POWER", "Power"}};
    for (auto& m : map) if (t == m[0]) return m[1];
    return nullptr;
}

int layer_from_node(const Node& n, bool v1, ProtoLayer& l, std::string& err) {
    for (const Field& f : n.fields) {
        if (f.key == "name") l.name = f.scalar;
        else if (f.key == "type") l.type = f.scalar;
        else if (f.key == "bottom") l.bottoms.push_back(f.scalar);
        else if (f.key == "top") l.tops.push_back(f.scalar);
    }
    if (v1) {
        const char* t = upgrade_v1_type(l.type);
        if (!t) { err = "layer " + l.name + ": legacy layer type " + l.type + " is not supported"; return -1; }
        l.type = t;
    }
    if (const Field* f = find(n, "convolution_param")) {
        if (!f->msg) { err = "layer " + l.name + ": convolution_param is not a message"; return -1; }
        const Node& c = *f->msg;
        int kh = 0, kw = 0, ph = -1, pw = -1, sh = 0, sw = 0, group = 1, dil = 1;
        if (!get_int(c, "num_output", l.num_output, err) || !get_int(c, "kernel_size", l.kernel, err) || !get_int(c, "pad", l.pad, err) ||
            !get_int(c, "stride", l.stride, err) || !get_int(c, "kernel_h", kh, err) || !get_int(c, "kernel_w", kw, err) ||
            !get_int(c, "pad_h", ph, err) || !get_int(c, "pad_w", pw, err) || !get_int(c, "stride_h", sh, err) ||
            !get_int(c, "stride_w", sw, err) || !get_int(c, "group", group, err) || !get_int(c, "dilation", dil, err)) {
            err = "layer " + l.name + ": " + err;
            return -1;
        }
        if (kh || kw) { if (kh != kw) { err = "layer " + l.name + ": non-square kernels are not supported"; return -1; } l.kernel = kh; }
        if (ph >= 0 || pw >= 0) { if (ph != pw) { err = "layer " + l.name + ": asymmetric padding is not supported"; return -1; } l.pad = ph; }
        if (sh || sw) { if (sh != sw) { err = "layer " + l.name + ": anisotropic stride is not supported"; return -1; } l.stride = sh; }
        if (group != 1 || dil != 1) { err = "layer " + l.name + ": group / dilation are not supported"; return -1; }
        if (const Field* b = find(c, "bias_term"))
            if (b->scalar == "false" || b->scalar == "0") { err = "layer " + l.name + ": bias_term false is not supported"; return -1; }
    }
    if (const Field* f = find(n, "pooling_param")) {
        if (!f->msg) { err = "layer " + l.name + ": pooling_param is not a message"; return -1; }
        const Node& c = *f->msg;
        l.stride = 1; l.pad = 0;
        if (!get_int(c, "kernel_size", l.kernel, err) || !get_int(c, "stride", l.stride, err) || !get_int(c, "pad", l.pad, err)) {
            err = "layer " + l.name + ": " + err;
            return -1;
        }
        if (const Field* p = find(c, "pool")) l.pool_method = (p->scalar == "MAX" || p->scalar == "0") ? 0 : (p->scalar == "AVE" || p->scalar == "1") ? 1 : 2;
    }
    if (const Field* f = find(n, "concat_param")) {
        if (f->msg && (!get_int(*f->msg, "axis", l.concat_axis, err) || !get_int(*f->msg, "concat_dim", l.concat_axis, err))) {
            err = "layer " + l.name + ": " + err;
            return -1;
        }
    }
    if (const Field* f = find(n, "nms_param")) {
        if (f->msg && (!get_float(*f->msg, "threshold", l.nms_threshold, err) || !get_int(*f->msg, "max_peaks", l.nms_max_peaks, err) ||
                       !get_int(*f->msg, "num_parts", l.nms_num_parts, err))) {
            err = "layer " + l.name + ": " + err;
            return -1;
        }
    }
    if (const Field* f = find(n, "imresize_param")) {
        if (f->msg && (!get_float(*f->msg, "factor", l.resize_factor, err) || !get_float(*f->msg, "start_scale", l.resize_start_scale, err) ||
                       !get_float(*f->msg, "scale_gap", l.resize_scale_gap, err))) {
            err = "layer " + l.name + ": " + err;
            return -1;
        }
    }
    return 0;
}

}  // namespace

int parse_prototxt_text(const std::string& text, NetDef& out, std::string& err) {
    Node root;
    Lexer lx(text);
    if (parse_node(lx, root, 0, 0, err)) return -1;
    out = NetDef();
    for (const Field& f : root.fields) {
        if (f.key == "name") out.name = f.scalar;
        else if (f.key == "input") out.inputs.push_back(f.scalar);
        else if (f.key == "input_dim") out.input_dims.push_back(atoi(f.scalar.c_str()));
        else if (f.key == "input_shape" && f.msg) {
            for (const Field& d : f.msg->fields) if (d.key == "dim") out.input_dims.push_back(atoi(d.scalar.c_str()));
        } else if ((f.key == "layer" || f.key == "layers") && f.msg) {
            ProtoLayer l;
            if (layer_from_node(*f.msg, f.key == "layers", l, err)) return -1;
            // NetStateRule: a deploy net is instantiated in phase TEST (rtpose.cpp:183); layers restricted to TRAIN are dropped
            bool skip = false;
            for (const Field& r : f.msg->fields)
                if (r.key == "include" && r.msg) if (const Field* ph = find(*r.msg, "phase")) if (ph->scalar == "TRAIN") skip = true;
            if (!skip) out.layers.push_back(l);
        }
    }
    if (out.layers.empty()) { err = "no layers found"; return -1; }
    return 0;
}

int parse_prototxt_file(const char* path, NetDef& out, std::string& err) {
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return -1; }
    std::string text;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) {
        text.append(buf, n);
        if (text.size() > (64u << 20)) { fclose(f); err = "prototxt larger than 64 MB"; return -1; }
    }
    fclose(f);
    return parse_prototxt_text(text, out, err);
}

// ---------------------------------------------------------------------------------------------
// the default deploy graphs in code, layer order as in the prototxts (L1 / L2 branches interleaved)
// ---------------------------------------------------------------------------------------------
NetDef builtin_netdef(int model, int stages) {
    const ModelTables& mt = model_tables(model);
    const int cL1 = 2 * mt.num_limbs, cL2 = mt.num_parts + 1;
    NetDef n;
    n.inputs = {"image"};
    n.input_dims = {1, 3, 368, 368};
    auto conv = [&](const std::string& name, const std::string& bottom, int cout, int k, const std::string& relu) {
        ProtoLayer l;
        l.name = name; l.type = "Convolution"; l.bottoms = {bottom}; l.tops = {name};
        l.num_output = cout; l.kernel = k; l.pad = k / 2; l.stride = 1;
        n.layers.push_back(l);
        if (!relu.empty()) {
            ProtoLayer r;
            r.name = relu; r.type = "ReLU"; r.bottoms = {name}; r.tops = {name};
            n.layers.push_back(r);
        }
    };
    char nm[64], rl[64];
    std::string prev = "image";
    const int vgg[4][2] = {{64, 2}, {128, 2}, {256, 4}, {512, 2}};
    for (int b = 0; b < 4; b++) {
        for (int i = 1; i <= vgg[b][1]; i++) {
            snprintf(nm, 64, "conv%d_%d", b + 1, i);
            snprintf(rl, 64, "relu%d_%d", b + 1, i);
            conv(nm, prev, vgg[b][0], 3, rl);
            prev = nm;
        }
        if (b < 3) {
            snprintf(nm, 64, "pool%d_stage1", b + 1);
            ProtoLayer l;
            l.name = nm; l.type = "Pooling"; l.bottoms = {prev}; l.tops = {nm}; l.kernel = 2; l.stride = 2; l.pad = 0; l.pool_method = 0;
            n.layers.push_back(l);
            prev = nm;
        }
    }
    conv("conv4_3_CPM", prev, 256, 3, "relu4_3_CPM");
    conv("conv4_4_CPM", "conv4_3_CPM", 128, 3, "relu4_4_CPM");
    std::string p1 = "conv4_4_CPM", p2 = "conv4_4_CPM";
    for (int i = 1; i <= 5; i++)
        for (int br = 1; br <= 2; br++) {
            snprintf(nm, 64, "conv5_%d_CPM_L%d", i, br);
            snprintf(rl, 64, "relu5_%d_CPM_L%d", i, br);
            std::string& p = br == 1 ? p1 : p2;
            if (i <= 3) conv(nm, p, 128, 3, rl);
            else if (i == 4) conv(nm, p, 512, 1, rl);
            else conv(nm, p, br == 1 ? cL1 : cL2, 1, "");
            p = nm;
        }
    for (int s = 2; s <= stages; s++) {
        snprintf(nm, 64, "concat_stage%d", s);
        ProtoLayer c;
        c.name = nm; c.type = "Concat"; c.bottoms = {p1, p2, "conv4_4_CPM"}; c.tops = {nm};
        n.layers.push_back(c);
        p1 = p2 = nm;
        for (int i = 1; i <= 7; i++)
            for (int br = 1; br <= 2; br++) {
                snprintf(nm, 64, "Mconv%d_stage%d_L%d", i, s, br);
                snprintf(rl, 64, "Mrelu%d_stage%d_L%d", i, s, br);
                std::string& p = br == 1 ? p1 : p2;
                if (i <= 5) conv(nm, p, 128, 7, rl);
                else if (i == 6) conv(nm, p, 128, 1, rl);
                else conv(nm, p, br == 1 ? cL1 : cL2, 1, "");
                p = nm;
            }
    }
    ProtoLayer c7;
    c7.name = "concat_stage7"; c7.type = "Concat"; c7.bottoms = {p2, p1}; c7.tops = {"concat_stage7"};   // [L2 | L1], prototxt :2966-2975
    n.layers.push_back(c7);
    ProtoLayer r;
    r.name = "resize"; r.type = "ImResize"; r.bottoms = {"concat_stage7"}; r.tops = {"resized_map"};
    r.resize_factor = 8; r.resize_start_scale = model == PE_MODEL_MPI_15 ? 0.9f : 1.f; r.resize_scale_gap = model == PE_MODEL_MPI_15 ? 0.1f : 0.3f;
    n.layers.push_back(r);
    ProtoLayer m;
    m.name = "nms"; m.type = "Nms"; m.bottoms = {"resized_map"}; m.tops = {"joints"};
    m.nms_threshold = model == PE_MODEL_MPI_15 ? 0.6f : 0.05f; m.nms_max_peaks = mt.max_peaks; m.nms_num_parts = mt.num_parts;
    n.layers.push_back(m);
    return n;
}

}  // namespace pe
