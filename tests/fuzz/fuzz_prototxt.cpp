// Mutation fuzzer for the deploy-prototxt reader and the plan builder (csrc/prototxt.cpp, csrc/plan.cpp): --caffeproto is a user file.
// Built with -fsanitize=address,undefined by tests/test_prototxt.py::test_prototxt_reader_survives_corrupt_files; any finding aborts.
// usage: fuzz_prototxt <iterations> <file.prototxt>...
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.h"
#include "prototxt.h"

namespace pe { int build_plan_from_net(const NetDef& net, int kp_input, int cpad, NetPlan& p, std::string& err); }

static std::string slurp(const char* p) {
    std::string d;
    FILE* f = fopen(p, "rb");
    if (!f) return d;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) d.append(buf, n);
    fclose(f);
    return d;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const int iters = atoi(argv[1]);
    uint64_t s = 1234567;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); };
    static const char* const kTokens[] = {"{", "}", ":", "\"", "layer", "layers", "bottom: \"", "top: \"", "num_output: ", "kernel_size: ", "pad: ",
                                          "stride: ", "-1", "0", "2147483647", "4294967296", "99999999999999999999", "1e39", "nan", "#", "\\", "\n",
                                          "type: \"Convolution\"", "type: CONVOLUTION", "type: \"Concat\"", "type: \"Pooling\"", "type: \"Nms\"",
                                          "type: \"ImResize\"", "input_dim: ", "nms_param {", "imresize_param {", "convolution_param {", "axis: "};
    long accepted = 0, planned = 0, rejected = 0;
    for (int a = 2; a < argc; a++) {
        const std::string base = slurp(argv[a]);
        if (base.size() < 64) return 2;
        std::vector<size_t> line_start{0};
        for (size_t i = 0; i + 1 < base.size(); i++) if (base[i] == '\n') line_start.push_back(i + 1);
        for (int it = 0; it < iters; it++) {
            std::string d = base;
            const int nm = 1 + rnd() % 5;
            for (int k = 0; k < nm && !d.empty(); k++) {
                const int kind = rnd() % 7;
                const size_t p = rnd() % d.size();
                if (kind == 0) d[p] = (char)rnd();
                else if (kind == 1) d.resize(p);                                             // cut off
                else if (kind == 2) d.insert(p, kTokens[rnd() % (sizeof kTokens / sizeof *kTokens)]);
                else if (kind == 3) {                                                        // drop a line
                    const size_t e = d.find('\n', p);
                    d.erase(p, e == std::string::npos ? std::string::npos : e - p + 1);
                } else if (kind == 4) {                                                      // repeat a block of the original somewhere else
                    const size_t b = line_start[rnd() % line_start.size()], len = std::min<size_t>(base.size() - b, 1 + rnd() % 600);
                    d.insert(std::min(p, d.size()), base, b, len);
                } else if (kind == 5) {                                                      // change a number
                    size_t q = d.find_first_of("0123456789", p);
                    if (q != std::string::npos) { size_t e = q; while (e < d.size() && isdigit((unsigned char)d[e])) e++; d.replace(q, e - q, kTokens[12 + rnd() % 7]); }
                } else {                                                                     // rename a blob reference
                    size_t q = d.find('"', p);
                    if (q != std::string::npos && q + 2 < d.size()) d[q + 1] = (char)('a' + rnd() % 26);
                }
            }
            pe::NetDef net;
            std::string err;
            if (pe::parse_prototxt_text(d, net, err)) { rejected++; if (err.empty()) { printf("rejected without a message\n"); return 3; } continue; }
            accepted++;
            pe::NetPlan plan;
            err.clear();
            if (pe::build_plan_from_net(net, 64, 64, plan, err) == 0) planned++;
            else if (err.empty()) { printf("plan refused without a message\n"); return 3; }
        }
    }
    printf("parsed %ld planned %ld rejected %ld\n", accepted, planned, rejected);
    return 0;
}
