// PNG decoder for --image_dir: the reference lists .jpg / .png / .bmp files (examples/rtpose/rtpose.cpp:1743) and reads
// them with cv::imread(path, CV_LOAD_IMAGE_COLOR) (:302-391), i.e. libpng + OpenCV's 8-bit BGR conversion.  No libpng /
// zlib headers for C++ in this image, so both layers are written out here:
//   * inflate (RFC 1951: stored, fixed and dynamic Huffman blocks) inside the zlib container (RFC 1950);
//   * PNG (ISO/IEC 15948): IHDR / PLTE / IDAT / IEND, the five scanline filters, Adam7 interlacing, colour types
//     0, 2, 3, 4, 6 at bit depths 1-16;
//   * conversion to what cv::imread(IMREAD_COLOR) returns: 16-bit samples keep their high byte (png_set_strip_16), grey
//     1/2/4-bit samples are scaled to 0..255, palette indices are expanded, grey becomes B=G=R, alpha and tRNS are dropped.
// Lossless, so "same pixels as the reference" is a property of the format; tests/test_abi.py checks it against cv2 for
// every colour type / depth cv2 can write plus hand-made palette, interlaced and low-bit-depth files.  CRCs and the
// Adler-32 are not verified (libpng would reject a corrupt file; here it decodes or fails on structure).  Host code.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/poseengine.h"

namespace {

struct Bits {
    const uint8_t* p; const uint8_t* end;
    uint32_t acc = 0;
    int n = 0;
    bool overrun = false;
    inline int get(int k) {   // k <= 16, LSB first
        while (n < k) {
            uint32_t b = 0;
            if (p < end) b = *p++; else overrun = true;
            acc |= b << n;
            n += 8;
        }
        const int v = (int)(acc & ((1u << k) - 1u));
        acc >>= k;
        n -= k;
        return v;
    }
};

struct Huffman {
    uint16_t count[16];
    uint16_t symbol[288];
    bool build(const uint8_t* lengths, int n) {   // canonical code from code lengths (RFC 1951 3.2.2)
        memset(count, 0, sizeof count);
        for (int i = 0; i < n; i++) count[lengths[i]]++;
        if (count[0] == n) return true;           // no codes: legal for an unused distance tree
        int left = 1;
        for (int l = 1; l < 16; l++) { left <<= 1; left -= count[l]; if (left < 0) return false; }
        uint16_t offs[16];
        offs[1] = 0;
        for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
        for (int i = 0; i < n; i++) if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
        return true;
    }
    inline int decode(Bits& b) const {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l < 16; l++) {
            code |= b.get(1);
            const int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
        }
        return -1;
    }
};

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

bool inflate_codes(Bits& b, const Huffman& lit, const Huffman& dist, std::vector<uint8_t>& out, size_t limit) {
    while (true) {
        int sym = lit.decode(b);
        if (sym < 0 || b.overrun) return false;
        if (sym < 256) { out.push_back((uint8_t)sym); if (out.size() > limit) return false; continue; }
        if (sym == 256) return true;
        sym -= 257;
        if (sym >= 29) return false;
        const int len = kLenBase[sym] + b.get(kLenExtra[sym]);
        const int ds = dist.decode(b);
        if (ds < 0 || ds >= 30) return false;
        const size_t d = (size_t)kDistBase[ds] + (size_t)b.get(kDistExtra[ds]);
        if (d > out.size() || out.size() + (size_t)len > limit) return false;
        const size_t start = out.size() - d;
        for (int i = 0; i < len; i++) out.push_back(out[start + i]);   // may overlap: byte by byte
    }
}

bool inflate_zlib(const uint8_t* data, size_t size, std::vector<uint8_t>& out, size_t limit) {
    if (size < 2 || (data[0] & 15) != 8 || ((data[0] << 8) | data[1]) % 31 != 0 || (data[1] & 0x20)) return false;
    Bits b;
    b.p = data + 2; b.end = data + size;
    out.clear();
    out.reserve(limit);
    int last;
    do {
        last = b.get(1);
        const int type = b.get(2);
        if (type == 0) {
            b.acc = 0; b.n = 0;   // to the byte boundary
            if (b.end - b.p < 4) return false;
            const int len = b.p[0] | (b.p[1] << 8), nlen = b.p[2] | (b.p[3] << 8);
            b.p += 4;
            if ((len ^ 0xFFFF) != nlen || b.end - b.p < len) return false;
            out.insert(out.end(), b.p, b.p + len);
            b.p += len;
        } else if (type == 1) {
            uint8_t l[288];
            for (int i = 0; i < 144; i++) l[i] = 8;
            for (int i = 144; i < 256; i++) l[i] = 9;
            for (int i = 256; i < 280; i++) l[i] = 7;
            for (int i = 280; i < 288; i++) l[i] = 8;
            uint8_t d[30];
            memset(d, 5, sizeof d);
            Huffman lit, dist;
            lit.build(l, 288); dist.build(d, 30);
            if (!inflate_codes(b, lit, dist, out, limit)) return false;
        } else if (type == 2) {
            const int nlen = b.get(5) + 257, ndist = b.get(5) + 1, ncode = b.get(4) + 4;
            if (nlen > 286 || ndist > 30) return false;
            uint8_t cl[19] = {0};
            for (int i = 0; i < ncode; i++) cl[kClOrder[i]] = (uint8_t)b.get(3);
            Huffman clh;
            if (!clh.build(cl, 19)) return false;
            uint8_t lengths[320];
            int idx = 0;
            while (idx < nlen + ndist) {
                const int sym = clh.decode(b);
                if (sym < 0 || b.overrun) return false;
                if (sym < 16) { lengths[idx++] = (uint8_t)sym; continue; }
                int rep, val = 0;
                if (sym == 16) { if (idx == 0) return false; val = lengths[idx - 1]; rep = 3 + b.get(2); }
                else if (sym == 17) rep = 3 + b.get(3);
                else rep = 11 + b.get(7);
                if (idx + rep > nlen + ndist) return false;
                while (rep--) lengths[idx++] = (uint8_t)val;
            }
            if (lengths[256] == 0) return false;
            Huffman lit, dist;
            if (!lit.build(lengths, nlen) || !dist.build(lengths + nlen, ndist)) return false;
            if (!inflate_codes(b, lit, dist, out, limit)) return false;
        } else {
            return false;
        }
        if (out.size() > limit || b.overrun) return false;
    } while (!last);
    return true;
}

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// undo the scanline filters of one (sub-)image in place; rows = h x (1 + stride) bytes
bool unfilter(uint8_t* d, int h, size_t stride, int bpp) {
    std::vector<uint8_t> zero(stride, 0);
    const uint8_t* prev = zero.data();
    for (int y = 0; y < h; y++) {
        uint8_t* row = d + (size_t)y * (stride + 1);
        const int f = row[0];
        uint8_t* cur = row + 1;
        switch (f) {
            case 0: break;
            case 1: for (size_t i = bpp; i < stride; i++) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]); break;
            case 2: for (size_t i = 0; i < stride; i++) cur[i] = (uint8_t)(cur[i] + prev[i]); break;
            case 3:
                for (size_t i = 0; i < stride; i++) cur[i] = (uint8_t)(cur[i] + (((i >= (size_t)bpp ? cur[i - bpp] : 0) + prev[i]) >> 1));
                break;
            case 4:
                for (size_t i = 0; i < stride; i++)
                    cur[i] = (uint8_t)(cur[i] + paeth(i >= (size_t)bpp ? cur[i - bpp] : 0, prev[i], i >= (size_t)bpp ? prev[i - bpp] : 0));
                break;
            default: return false;
        }
        prev = cur;
    }
    return true;
}

}  // namespace

// PNG bytes -> uint8 BGR HWC as cv::imread(IMREAD_COLOR).  Same contract as pe_decode_jpeg.
extern "C" int pe_decode_png(const uint8_t* data, long long size, int* w, int* h, uint8_t* bgr, long long cap) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (!data || size < 8 + 25 || memcmp(data, sig, 8) != 0) return -1;
    const uint8_t* p = data + 8;
    const uint8_t* end = data + size;
    int W = 0, H = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    bool have_ihdr = false, done = false;
    while (!done && p + 12 <= end) {
        const uint32_t len = be32(p);
        const uint8_t* type = p + 4;
        const uint8_t* body = p + 8;
        if (len > (uint32_t)(end - body) || (size_t)(end - body) < (size_t)len + 4) return -1;
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) return -1;
            W = (int)be32(body); H = (int)be32(body + 4);
            depth = body[8]; ctype = body[9]; interlace = body[12];
            if (W <= 0 || H <= 0 || body[10] != 0 || body[11] != 0 || interlace > 1) return -1;
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(body, body + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!memcmp(type, "IEND", 4)) {
            done = true;
        }
        p = body + len + 4;
    }
    if (!have_ihdr || idat.empty()) return -1;
    int channels;
    switch (ctype) {
        case 0: channels = 1; if (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16) return -1; break;
        case 2: channels = 3; if (depth != 8 && depth != 16) return -1; break;
        case 3: channels = 1; if (depth != 1 && depth != 2 && depth != 4 && depth != 8) return -1; if (plte.size() < 3) return -1; break;
        case 4: channels = 2; if (depth != 8 && depth != 16) return -1; break;
        case 6: channels = 4; if (depth != 8 && depth != 16) return -1; break;
        default: return -1;
    }
    if (w) *w = W;
    if (h) *h = H;
    if (!bgr) return 0;
    if (cap < (long long)W * H * 3) return -1;
    const int bits_pp = channels * depth;
    const int bpp = bits_pp >= 8 ? bits_pp / 8 : 1;
    auto row_bytes = [&](int pw) { return ((size_t)pw * bits_pp + 7) / 8; };
    // expected size of the filtered stream
    static const int xs[7] = {0, 4, 0, 2, 0, 1, 0}, ys[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
    size_t total = 0;
    if (!interlace) total = (size_t)H * (row_bytes(W) + 1);
    else
        for (int k = 0; k < 7; k++) {
            const int pw = (W - xs[k] + dx[k] - 1) / dx[k], ph = (H - ys[k] + dy[k] - 1) / dy[k];
            if (pw > 0 && ph > 0) total += (size_t)ph * (row_bytes(pw) + 1);
        }
    std::vector<uint8_t> raw;
    if (!inflate_zlib(idat.data(), idat.size(), raw, total) || raw.size() < total) return -1;
    // one sample (8-bit view, as cv::imread gives it) of channel c of pixel x in an unfiltered row
    auto sample8 = [&](const uint8_t* row, int x, int c) -> int {
        if (depth == 8) return row[(size_t)x * channels + c];
        if (depth == 16) return row[((size_t)x * channels + c) * 2];   // high byte: png_set_strip_16
        const int per = 8 / depth, v = (row[x / per] >> ((per - 1 - x % per) * depth)) & ((1 << depth) - 1);
        return v;
    };
    auto put = [&](const uint8_t* row, int x, uint8_t* o) {
        if (ctype == 3) {
            const size_t i = (size_t)sample8(row, x, 0) * 3;
            if (i + 2 < plte.size()) { o[2] = plte[i]; o[1] = plte[i + 1]; o[0] = plte[i + 2]; }
            else o[0] = o[1] = o[2] = 0;
        } else if (ctype == 0 || ctype == 4) {
            int g = sample8(row, x, 0);
            if (depth < 8) g = g * 255 / ((1 << depth) - 1);   // png_set_expand_gray_1_2_4_to_8
            o[0] = o[1] = o[2] = (uint8_t)g;
        } else {
            o[2] = (uint8_t)sample8(row, x, 0); o[1] = (uint8_t)sample8(row, x, 1); o[0] = (uint8_t)sample8(row, x, 2);
        }
    };
    uint8_t* d = raw.data();
    if (!interlace) {
        const size_t stride = row_bytes(W);
        if (!unfilter(d, H, stride, bpp)) return -1;
        for (int y = 0; y < H; y++) {
            const uint8_t* row = d + (size_t)y * (stride + 1) + 1;
            for (int x = 0; x < W; x++) put(row, x, bgr + ((size_t)y * W + x) * 3);
        }
        return 0;
    }
    for (int k = 0; k < 7; k++) {   // Adam7
        const int pw = (W - xs[k] + dx[k] - 1) / dx[k], ph = (H - ys[k] + dy[k] - 1) / dy[k];
        if (pw <= 0 || ph <= 0) continue;
        const size_t stride = row_bytes(pw);
        if (!unfilter(d, ph, stride, bpp)) return -1;
        for (int y = 0; y < ph; y++) {
            const uint8_t* row = d + (size_t)y * (stride + 1) + 1;
            for (int x = 0; x < pw; x++) put(row, x, bgr + ((size_t)(ys[k] + y * dy[k]) * W + xs[k] + x * dx[k]) * 3);
        }
        d += (size_t)ph * (stride + 1);
    }
    return 0;
}
