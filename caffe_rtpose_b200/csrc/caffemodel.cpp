// .caffemodel reader: Net::CopyTrainedLayersFrom for binary NetParameter files (src/caffe/net.cpp:750-803,
// src/caffe/util/io.cpp ReadProtoFromBinaryFile) without protobuf/protoc: a minimal protobuf WIRE-FORMAT walker
// for exactly the fields the path needs (src/caffe/proto/caffe.proto):
//   NetParameter     layer = 100 (LayerParameter), layers = 2 (V1LayerParameter, legacy files)
//   LayerParameter   name = 1, type = 2, blobs = 7          V1LayerParameter  name = 4, type = 5 (enum), blobs = 6
//   BlobProto        data = 5 (packed or repeated float), shape = 7 (BlobShape.dim = 1), num..width = 1..4 (legacy)
// Semantics kept from the reference: layers are matched BY NAME, source layers unknown to the net are ignored,
// a blob-count or shape mismatch is an error (CHECK / LOG(FATAL) there, PE_ERR_INVALID here); legacy 4-D dims and
// BlobShape are both accepted (blob.cpp:448-470).
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/poseengine.h"

namespace {

struct Blob { std::vector<float> data; std::vector<long long> shape; };
struct Layer { std::string name, type; std::vector<Blob> blobs; };

struct Reader {
    const uint8_t* p; const uint8_t* end; bool ok = true;
    Reader(const uint8_t* b, size_t n) : p(b), end(b + n) {}
    bool done() const { return p >= end || !ok; }
    uint64_t varint() {
        uint64_t v = 0; int shift = 0;
        while (p < end && shift < 64) {
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
        }
        ok = false; return 0;
    }
    Reader sub() {   // length-delimited payload
        const uint64_t n = varint();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return Reader(p, 0); }
        Reader r(p, (size_t)n); p += n; return r;
    }
    void skip(int wt) {
        if (wt == 0) varint();
        else if (wt == 1) { if (end - p < 8) ok = false; else p += 8; }
        else if (wt == 2) sub();
        else if (wt == 5) { if (end - p < 4) ok = false; else p += 4; }
        else ok = false;
    }
};

bool parse_blob(Reader r, Blob& b) {
    long long legacy[4] = {0, 0, 0, 0}; bool have_legacy = false;
    while (!r.done()) {
        const uint64_t key = r.varint(); const int f = (int)(key >> 3), wt = (int)(key & 7);
        if (f == 5 && wt == 2) {            // packed floats
            Reader d = r.sub();
            const size_t n = (size_t)(d.end - d.p) / 4;
            const size_t old = b.data.size();
            b.data.resize(old + n);
            if (n) memcpy(b.data.data() + old, d.p, n * 4);
        } else if (f == 5 && wt == 5) {     // unpacked repeated float
            if (r.end - r.p < 4) return false;
            float v; memcpy(&v, r.p, 4); r.p += 4; b.data.push_back(v);
        } else if (f == 7 && wt == 2) {     // BlobShape
            Reader s = r.sub();
            while (!s.done()) {
                const uint64_t k2 = s.varint();
                if ((k2 >> 3) == 1 && (k2 & 7) == 2) { Reader d = s.sub(); while (!d.done()) b.shape.push_back((long long)d.varint()); }
                else if ((k2 >> 3) == 1 && (k2 & 7) == 0) b.shape.push_back((long long)s.varint());
                else s.skip((int)(k2 & 7));
            }
            if (!s.ok) return false;
        } else if (f >= 1 && f <= 4 && wt == 0) { legacy[f - 1] = (long long)r.varint(); have_legacy = true; }
        else r.skip(wt);
    }
    if (b.shape.empty() && have_legacy) b.shape.assign(legacy, legacy + 4);
    return r.ok;
}

bool parse_layer(Reader r, bool v1, Layer& l) {
    const int f_name = v1 ? 4 : 1, f_type = v1 ? 5 : 2, f_blobs = v1 ? 6 : 7;
    while (!r.done()) {
        const uint64_t key = r.varint(); const int f = (int)(key >> 3), wt = (int)(key & 7);
        if (f == f_name && wt == 2) { Reader s = r.sub(); l.name.assign((const char*)s.p, (size_t)(s.end - s.p)); }
        else if (f == f_type && wt == 2) { Reader s = r.sub(); l.type.assign((const char*)s.p, (size_t)(s.end - s.p)); }
        else if (f == f_type && wt == 0) { l.type = "V1:" + std::to_string((long long)r.varint()); }
        else if (f == f_blobs && wt == 2) { Blob b; if (!parse_blob(r.sub(), b)) return false; l.blobs.push_back(std::move(b)); }
        else r.skip(wt);
    }
    return r.ok;
}

}  // namespace

struct pe_caffemodel { std::vector<Layer> layers; };

static thread_local std::string g_cm_error;
extern "C" const char* pe_caffemodel_last_error(void) { return g_cm_error.c_str(); }

extern "C" int pe_caffemodel_open(const char* path, pe_caffemodel** out) {
    if (!path || !out) return PE_ERR_INVALID;
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) { g_cm_error = std::string("cannot open ") + path; return PE_ERR_IO; }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf((size_t)(n > 0 ? n : 0));
    const bool rd = n >= 0 && fread(buf.data(), 1, buf.size(), f) == buf.size();
    fclose(f);
    if (!rd) { g_cm_error = std::string("cannot read ") + path; return PE_ERR_IO; }
    pe_caffemodel* m = new pe_caffemodel();
    Reader r(buf.data(), buf.size());
    while (!r.done()) {
        const uint64_t key = r.varint(); const int fld = (int)(key >> 3), wt = (int)(key & 7);
        if ((fld == 100 || fld == 2) && wt == 2) {
            Layer l;
            if (!parse_layer(r.sub(), fld == 2, l)) { r.ok = false; break; }
            m->layers.push_back(std::move(l));
        } else r.skip(wt);
    }
    if (!r.ok) { delete m; g_cm_error = std::string(path) + ": not a valid binary NetParameter"; return PE_ERR_IO; }
    *out = m;
    return PE_OK;
}
extern "C" void pe_caffemodel_close(pe_caffemodel* m) { delete m; }
extern "C" int pe_caffemodel_num_layers(const pe_caffemodel* m) { return m ? (int)m->layers.size() : 0; }
extern "C" int pe_caffemodel_layer(const pe_caffemodel* m, int idx, char* name64, char* type32, int* num_blobs) {
    if (!m || idx < 0 || idx >= (int)m->layers.size()) return PE_ERR_INVALID;
    const Layer& l = m->layers[idx];
    if (name64) snprintf(name64, 64, "%s", l.name.c_str());
    if (type32) snprintf(type32, 32, "%s", l.type.c_str());
    if (num_blobs) *num_blobs = (int)l.blobs.size();
    return PE_OK;
}
extern "C" int pe_caffemodel_blob(const pe_caffemodel* m, int layer, int blob, const float** data, size_t* count, int* ndim,
                                  long long* dims8) {
    if (!m || layer < 0 || layer >= (int)m->layers.size()) return PE_ERR_INVALID;
    const Layer& l = m->layers[layer];
    if (blob < 0 || blob >= (int)l.blobs.size()) return PE_ERR_INVALID;
    const Blob& b = l.blobs[blob];
    if (data) *data = b.data.data();
    if (count) *count = b.data.size();
    if (ndim) *ndim = (int)b.shape.size();
    if (dims8) for (size_t i = 0; i < b.shape.size() && i < 8; i++) dims8[i] = b.shape[i];
    return PE_OK;
}

// Net::CopyTrainedLayersFrom(trained_filename)  (rtpose.cpp:184)
extern "C" int pe_load_caffemodel(pe_engine* e, const char* path) {
    if (!e || !path) return PE_ERR_INVALID;
    pe_caffemodel* m = nullptr;
    const int rc = pe_caffemodel_open(path, &m);
    if (rc) return rc;
    int status = PE_OK;
    for (const Layer& l : m->layers) {
        if (l.blobs.empty()) continue;                       // ReLU/Pooling/... carry no blobs
        int idx = -1, cout = 0, cin = 0, k = 0;
        char name[64];
        for (int i = 0; i < pe_num_conv_layers(e); i++) {
            pe_conv_layer_info(e, i, name, &cout, &cin, &k);
            if (l.name == name) { idx = i; break; }
        }
        if (idx < 0) continue;                               // "Ignoring source layer" (net.cpp:757-763)
        if (l.blobs.size() != 2) { g_cm_error = "layer " + l.name + ": incompatible number of blobs"; status = PE_ERR_INVALID; break; }
        status = pe_set_conv_weights(e, l.name.c_str(), l.blobs[0].data.data(), l.blobs[0].data.size(), l.blobs[1].data.data(),
                                     l.blobs[1].data.size());
        if (status) break;
        // shape check beyond the element count: (cout, cin, k, k) in either BlobShape or legacy form
        const std::vector<long long>& s = l.blobs[0].shape;
        if (s.size() == 4 && (s[0] != cout || s[1] != cin || s[2] != k || s[3] != k)) {
            g_cm_error = "layer " + l.name + ": weight shape mismatch"; status = PE_ERR_INVALID; break;
        }
    }
    pe_caffemodel_close(m);
    return status;
}
