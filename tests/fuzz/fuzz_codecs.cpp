// Mutation fuzzer for the parsers that read untrusted files: pe_decode_jpeg, pe_decode_png, the .caffemodel reader, the AVI reader.
// Built with -fsanitize=address,undefined by tests/test_abi.py::test_parsers_survive_corrupt_files; any finding aborts.
// usage: fuzz_codecs <iterations> <file>...   (.jpg / .png / .avi by signature, anything else is treated as a .caffemodel)
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "poseengine.h"

// the .caffemodel reader also holds pe_load_caffemodel, which needs these three engine entry points
extern "C" int pe_num_conv_layers(const pe_engine*) { return 0; }
extern "C" int pe_conv_layer_info(const pe_engine*, int, char*, int*, int*, int*) { return 1; }
extern "C" int pe_set_conv_weights(pe_engine*, const char*, const float*, size_t, const float*, size_t) { return 0; }

static std::vector<uint8_t> slurp(const char* p) {
    std::vector<uint8_t> d;
    FILE* f = fopen(p, "rb");
    if (!f) return d;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    d.resize(n > 0 ? (size_t)n : 0);
    if (fread(d.data(), 1, d.size(), f) != d.size()) d.clear();
    fclose(f);
    return d;
}

// Fixed regression inputs (exact-size heap buffers, so ASAN sees any byte read past the end).
static void fixed_cases() {
    // SOS segment whose length field says "no payload" at the very end of the file - the component count
    // used to be read before the length check (1-byte heap over-read).
    const uint8_t sos_trunc[] = {0xFF, 0xD8, 0xFF, 0xC0, 0x00, 0x0B, 0x08, 0x00, 0x08, 0x00, 0x08, 0x01, 0x01, 0x11, 0x00, 0xFF, 0xDA, 0x00, 0x02};
    std::vector<uint8_t> d(sos_trunc, sos_trunc + sizeof sos_trunc);
    int w = 0, h = 0;
    std::vector<uint8_t> px(8 * 8 * 3);
    if (pe_decode_jpeg(d.data(), (long long)d.size(), &w, &h, px.data(), (long long)px.size()) == 0) { printf("truncated SOS accepted\n"); exit(3); }
    for (size_t cut = 2; cut < d.size(); cut++) {   // every prefix, too
        std::vector<uint8_t> e(d.begin(), d.begin() + cut);
        pe_decode_jpeg(e.data(), (long long)e.size(), &w, &h, px.data(), (long long)px.size());
    }
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    fixed_cases();
    const int iters = atoi(argv[1]);
    uint64_t s = 987654321;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); };
    long ok = 0, rejected = 0;
    const std::string tmp = std::string(argv[2]) + ".mut";
    for (int a = 2; a < argc; a++) {
        const std::vector<uint8_t> base = slurp(argv[a]);
        if (base.size() < 16) return 2;
        const bool jpg = base[0] == 0xFF && base[1] == 0xD8, png = base[1] == 'P' && base[2] == 'N', avi = !memcmp(base.data(), "RIFF", 4);
        for (int it = 0; it < iters; it++) {
            std::vector<uint8_t> d = base;
            const int nm = 1 + rnd() % 6;
            for (int k = 0; k < nm; k++) {
                const int kind = rnd() % 4;
                if (kind == 0) d[rnd() % d.size()] = (uint8_t)rnd();
                else if (kind == 1) d[rnd() % d.size()] ^= (uint8_t)(1u << (rnd() % 8));
                else if (kind == 2 && d.size() > 16) d.resize(8 + rnd() % (d.size() - 8));
                else { const size_t p = rnd() % d.size(); d[p] = 0xFF; if (p + 1 < d.size()) d[p + 1] = (uint8_t)(0xC0 + rnd() % 0x20); }
            }
            if (jpg || png) {
                int w = 0, h = 0;
                auto dec = png ? pe_decode_png : pe_decode_jpeg;
                int rc = dec(d.data(), (long long)d.size(), &w, &h, nullptr, 0);
                if (rc == 0 && (long long)w * h <= 4000000) {
                    std::vector<uint8_t> px((size_t)w * h * 3);
                    rc = dec(d.data(), (long long)d.size(), &w, &h, px.data(), (long long)px.size());
                }
                (rc == 0 ? ok : rejected)++;
            } else if (avi) {
                FILE* f = fopen(tmp.c_str(), "wb");
                if (!f) return 2;
                fwrite(d.data(), 1, d.size(), f);
                fclose(f);
                pe_video* v = nullptr;
                if (pe_video_open(tmp.c_str(), &v) != 0) { rejected++; continue; }
                int w = 0, h = 0, n = 0;
                double fps = 0;
                char cc[5];
                pe_video_info(v, &w, &h, &fps, &n, cc);
                bool all = true;
                if ((long long)w * h <= 4000000) {
                    std::vector<uint8_t> px((size_t)w * h * 3);   // exact size: a row written past the frame is an ASAN finding
                    for (int i = 0; i < n && i < 64; i++) all = (pe_video_read(v, i, px.data(), (long long)px.size()) == 0) && all;
                }
                (all ? ok : rejected)++;
                pe_video_close(v);
            } else {
                FILE* f = fopen(tmp.c_str(), "wb");
                if (!f) return 2;
                fwrite(d.data(), 1, d.size(), f);
                fclose(f);
                pe_caffemodel* m = nullptr;
                if (pe_caffemodel_open(tmp.c_str(), &m) != 0) { rejected++; continue; }
                ok++;
                double acc = 0;
                for (int l = 0; l < pe_caffemodel_num_layers(m); l++) {
                    char name[64], type[32];
                    int nb = 0;
                    if (pe_caffemodel_layer(m, l, name, type, &nb)) continue;
                    for (int b = 0; b < nb; b++) {
                        const float* data = nullptr;
                        size_t cnt = 0;
                        int nd = 0;
                        long long dims[8];
                        if (pe_caffemodel_blob(m, l, b, &data, &cnt, &nd, dims) == 0 && data)
                            for (size_t i = 0; i < cnt; i++) acc += data[i];
                    }
                }
                if (acc == 12345.678) printf("!");
                pe_caffemodel_close(m);
            }
        }
    }
    remove(tmp.c_str());
    printf("accepted %ld rejected %ld\n", ok, rejected);
    return 0;
}
