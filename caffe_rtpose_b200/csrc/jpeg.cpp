// Baseline JPEG (JFIF) encoder for --write_frames: the reference writes the rendered frame with
// cv::imwrite(..., {CV_IMWRITE_JPEG_QUALITY, 98}) (examples/rtpose/rtpose.cpp:1363-1380); this image has no libjpeg /
// OpenCV for C++, so the encoder is written out here: ITU-T T.81 baseline sequential DCT, 8-bit, YCbCr 4:2:0 (the
// libjpeg default that cv::imwrite uses), Annex K quantisation tables scaled with libjpeg's quality rule, Annex K
// Huffman tables.  The byte stream is a standard JFIF file; it is NOT bit-identical to libjpeg's output (different DCT
// rounding), which no test of the reference depends on.  Host code, no GPU.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/poseengine.h"

#include "jpeg_tables.h"
namespace {
using namespace pe_jpeg;

const uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
// T.81 Annex K.1 (natural order)
const uint8_t kQLum[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                           18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kQChr[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                           99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
struct Huff { uint16_t code[256]; uint8_t len[256]; };
void build_huff(const uint8_t* bits, const uint8_t* vals, Huff& h) {   // T.81 Annex C
    memset(&h, 0, sizeof h);
    int k = 0;
    uint16_t code = 0;
    for (int l = 1; l <= 16; l++) {
        for (int i = 0; i < bits[l - 1]; i++) { h.code[vals[k]] = code++; h.len[vals[k]] = (uint8_t)l; k++; }
        code <<= 1;
    }
}

// Entropy-coded bytes go into a buffer that is grown once per MCU (six blocks x 64 coefficients x at most 26 bits, doubled by byte
// stuffing, stay below 4 KB), so put() never allocates; a Huffman code and the value bits that follow it leave in one call.
struct BitWriter {
    std::vector<uint8_t>& out;
    uint8_t* p = nullptr;
    uint64_t acc = 0;
    int n = 0;
    explicit BitWriter(std::vector<uint8_t>& o) : out(o) {}
    void begin_mcu() {
        const size_t used = out.size();
        if (out.capacity() - used < 4096) out.reserve(out.capacity() * 2 + 65536);
        out.resize(used + 4096);
        p = out.data() + used;
    }
    void end_mcu() { out.resize((size_t)(p - out.data())); }
    inline void put(uint32_t bits, int len) {   // len <= 32; at most 7 bits wait in acc
        acc = (acc << len) | (bits & (uint32_t)((1ull << len) - 1ull));
        n += len;
        while (n >= 8) {
            const uint8_t b = (uint8_t)(acc >> (n - 8));
            *p++ = b;
            if (b == 0xFF) *p++ = 0;   // byte stuffing
            n -= 8;
        }
    }
    void flush() { if (n) put(0x7F, 8 - n); }   // pad with ones
};

struct Basis { double t[8][8]; };   // t[x][u] = c(u) cos((2x + 1) u pi / 16): separable DCT-II, orthonormal JPEG scaling
const Basis& dct_basis() {
    static const Basis basis = [] {   // thread-safe one-time initialisation
        Basis b;
        for (int u = 0; u < 8; u++)
            for (int x = 0; x < 8; x++) b.t[x][u] = (u == 0 ? sqrt(0.125) : 0.5) * cos((2 * x + 1) * u * M_PI / 16.0);
        return b;
    }();
    return basis;
}
// Forward DCT + quantisation of one block: nat[i] = lrint(F[i] / q[i]), natural order.  Every DCT output is the sum over x (then y)
// ascending of basis * sample, started from zero - the order a scalar loop per output uses - but written with the output index
// innermost, so that the compiler keeps eight independent sums in vector registers; no fused multiply-add (the x86-64 baseline and
// target("avx2") have none), so the generic and the AVX2 instance produce the same bits.
static inline __attribute__((always_inline)) void fdct_quant_impl(const float* in, const uint8_t* q, int* nat, const double (*t)[8]) {
    double tmp[64], f[64];
    for (int y = 0; y < 8; y++) {
        double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int x = 0; x < 8; x++) {
            const double v = in[y * 8 + x];
            for (int u = 0; u < 8; u++) s[u] += t[x][u] * v;
        }
        for (int u = 0; u < 8; u++) tmp[y * 8 + u] = s[u];
    }
    for (int v = 0; v < 8; v++) {
        double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int y = 0; y < 8; y++) {
            const double c = t[y][v];
            for (int u = 0; u < 8; u++) s[u] += c * tmp[y * 8 + u];
        }
        for (int u = 0; u < 8; u++) f[v * 8 + u] = s[u];
    }
    for (int i = 0; i < 64; i++) f[i] = f[i] / q[i];
    for (int i = 0; i < 64; i++) nat[i] = (int)lrint(f[i]);
}
static void fdct_quant_generic(const float* in, const uint8_t* q, int* nat, const double (*t)[8]) { fdct_quant_impl(in, q, nat, t); }
#if defined(__x86_64__) && defined(__GNUC__) && !defined(PE_JPEG_NO_AVX2)   // (tests build the generic instance with this macro)
__attribute__((target("avx2"))) static void fdct_quant_avx2(const float* in, const uint8_t* q, int* nat, const double (*t)[8]) { fdct_quant_impl(in, q, nat, t); }
static const bool g_enc_avx2 = __builtin_cpu_supports("avx2");
#else
static void fdct_quant_avx2(const float* in, const uint8_t* q, int* nat, const double (*t)[8]) { fdct_quant_impl(in, q, nat, t); }
static const bool g_enc_avx2 = false;
#endif

inline int bit_size(int v) { const unsigned a = (unsigned)(v < 0 ? -v : v); return a ? 32 - __builtin_clz(a) : 0; }

void encode_block(const float* px, const uint8_t* q, int& dc_pred, const Huff& dc, const Huff& ac, BitWriter& bw) {
    int nat[64], z[64];
    (g_enc_avx2 ? fdct_quant_avx2 : fdct_quant_generic)(px, q, nat, dct_basis().t);
    for (int i = 0; i < 64; i++) {
        const int v = nat[kZigzag[i]];
        const int lo = i == 0 ? -1024 : -1023;   // 8-bit baseline coefficient range (T.81 F.1.2): DC diff fits 11 bits, AC 10
        z[i] = v < lo ? lo : (v > 1023 ? 1023 : v);
    }
    const int diff = z[0] - dc_pred;
    dc_pred = z[0];
    int s = bit_size(diff);
    bw.put(((uint32_t)dc.code[s] << s) | ((uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << s) - 1u)), dc.len[s] + s);
    int run = 0;
    for (int i = 1; i < 64; i++) {
        if (z[i] == 0) { run++; continue; }
        while (run > 15) { bw.put(ac.code[0xF0], ac.len[0xF0]); run -= 16; }
        s = bit_size(z[i]);
        const int sym = (run << 4) | s;
        bw.put(((uint32_t)ac.code[sym] << s) | ((uint32_t)(z[i] < 0 ? z[i] - 1 : z[i]) & ((1u << s) - 1u)), ac.len[sym] + s);
        run = 0;
    }
    if (run) bw.put(ac.code[0], ac.len[0]);   // EOB
}

void put16(std::vector<uint8_t>& o, int v) { o.push_back((uint8_t)(v >> 8)); o.push_back((uint8_t)v); }

}  // namespace

// BGR uint8 HWC image -> JFIF bytes.  Returns the length (writes if <= cap), or -1 for bad arguments.
extern "C" long long pe_encode_jpeg(const uint8_t* bgr, int w, int h, int quality, uint8_t* buf, long long cap) {
    if (!bgr || w <= 0 || h <= 0 || w > 65535 || h > 65535) return -1;
    quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;   // libjpeg jpeg_quality_scaling
    uint8_t ql[64], qc[64];
    for (int i = 0; i < 64; i++) {
        int a = (kQLum[i] * scale + 50) / 100, b = (kQChr[i] * scale + 50) / 100;
        ql[i] = (uint8_t)(a < 1 ? 1 : (a > 255 ? 255 : a));
        qc[i] = (uint8_t)(b < 1 ? 1 : (b > 255 ? 255 : b));
    }
    std::vector<uint8_t> o;
    o.reserve((size_t)w * h / 2 + 1024);
    o.push_back(0xFF); o.push_back(0xD8);                                           // SOI
    const uint8_t app0[] = {0xFF, 0xE0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    o.insert(o.end(), app0, app0 + sizeof app0);
    for (int t = 0; t < 2; t++) {                                                    // DQT (zigzag order)
        o.push_back(0xFF); o.push_back(0xDB); put16(o, 67); o.push_back((uint8_t)t);
        for (int i = 0; i < 64; i++) o.push_back((t ? qc : ql)[kZigzag[i]]);
    }
    o.push_back(0xFF); o.push_back(0xC0); put16(o, 17); o.push_back(8); put16(o, h); put16(o, w); o.push_back(3);   // SOF0
    o.push_back(1); o.push_back(0x22); o.push_back(0);                               // Y  2x2, table 0
    o.push_back(2); o.push_back(0x11); o.push_back(1);                               // Cb 1x1, table 1
    o.push_back(3); o.push_back(0x11); o.push_back(1);                               // Cr
    const uint8_t* bits[4] = {kDcLumBits, kAcLumBits, kDcChrBits, kAcChrBits};
    const uint8_t* vals[4] = {kDcVal, kAcLumVal, kDcVal, kAcChrVal};
    const int nval[4] = {12, 162, 12, 162};
    const uint8_t cls[4] = {0x00, 0x10, 0x01, 0x11};
    for (int t = 0; t < 4; t++) {                                                    // DHT
        o.push_back(0xFF); o.push_back(0xC4); put16(o, 19 + nval[t]); o.push_back(cls[t]);
        o.insert(o.end(), bits[t], bits[t] + 16);
        o.insert(o.end(), vals[t], vals[t] + nval[t]);
    }
    const uint8_t sos[] = {0xFF, 0xDA, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
    o.insert(o.end(), sos, sos + sizeof sos);
    Huff hdl, hal, hdc, hac;
    build_huff(kDcLumBits, kDcVal, hdl); build_huff(kAcLumBits, kAcLumVal, hal);
    build_huff(kDcChrBits, kDcVal, hdc); build_huff(kAcChrBits, kAcChrVal, hac);
    BitWriter bw(o);
    int pred[3] = {0, 0, 0};
    float Y[4][64], Cb[64], Cr[64], yy[16][16], cb[16][16], cr[16][16];
    for (int my = 0; my < h; my += 16)
        for (int mx = 0; mx < w; mx += 16) {
            for (int y = 0; y < 16; y++)
                for (int x = 0; x < 16; x++) {
                    const int sy = my + y < h ? my + y : h - 1, sx = mx + x < w ? mx + x : w - 1;   // edge replication
                    const uint8_t* p = bgr + ((size_t)sy * w + sx) * 3;
                    const float B = p[0], G = p[1], R = p[2];
                    yy[y][x] = 0.299f * R + 0.587f * G + 0.114f * B - 128.f;
                    cb[y][x] = -0.168736f * R - 0.331264f * G + 0.5f * B;
                    cr[y][x] = 0.5f * R - 0.418688f * G - 0.081312f * B;
                }
            for (int b = 0; b < 4; b++)
                for (int y = 0; y < 8; y++)
                    for (int x = 0; x < 8; x++) Y[b][y * 8 + x] = yy[(b >> 1) * 8 + y][(b & 1) * 8 + x];
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) {                                         // 2x2 box for the 4:2:0 chroma
                    Cb[y * 8 + x] = 0.25f * (cb[2 * y][2 * x] + cb[2 * y][2 * x + 1] + cb[2 * y + 1][2 * x] + cb[2 * y + 1][2 * x + 1]);
                    Cr[y * 8 + x] = 0.25f * (cr[2 * y][2 * x] + cr[2 * y][2 * x + 1] + cr[2 * y + 1][2 * x] + cr[2 * y + 1][2 * x + 1]);
                }
            bw.begin_mcu();
            for (int b = 0; b < 4; b++) encode_block(Y[b], ql, pred[0], hdl, hal, bw);
            encode_block(Cb, qc, pred[1], hdc, hac, bw);
            encode_block(Cr, qc, pred[2], hdc, hac, bw);
            bw.end_mcu();
        }
    bw.begin_mcu();
    bw.flush();
    bw.end_mcu();
    o.push_back(0xFF); o.push_back(0xD9);                                           // EOI
    if (buf && (long long)o.size() <= cap) memcpy(buf, o.data(), o.size());
    return (long long)o.size();
}
