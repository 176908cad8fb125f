// host/jpeg_decode.cpp - baseline JPEG decoder for bitmap textures.
//
// The reference decodes bitmap files through stb_image (scene/texture/bitmap.hpp:11-37: stbi_load, CMakeLists.txt:17-21 fetches
// nothings/stb `master`, unpinned and absent offline).  JPEG decoders are allowed to differ in the last bits (IDCT, chroma
// upsampling, colour conversion), and such differences reach the frame through every texel, so this decoder restates the
// ARITHMETIC stb_image's JPEG path is published to use - the 12-bit fixed-point "jidctint" inverse DCT with its +512 >> 10 column
// and +65536+(128<<17) >> 17 row roundings, the 3:1 / 9:3:3:1 triangle-filter chroma upsampling, the 20-bit fixed-point
// YCbCr -> RGB rows - around a plain bit-serial Huffman reader.  Pin: with this decoder the bitmap quadrant of the reference's
// published outputs/textures.png is reproduced exactly (tests/test_oracle_golden.py, tests/test_host.py); the file the reference
// ships (scenes/hw12/textures/dragon.jpg) is baseline, 4:4:4, no restart markers, so the upsampling and restart paths are
// exercised by synthetic files only.
//
// Supported: SOF0 / SOF1 with 8-bit samples, 1 or 3 components, sampling factors 1 or 2, restart intervals.  Progressive,
// arithmetic-coded, 12-bit and CMYK files are refused (RT_ERR_UNSUPPORTED).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "scene.hpp"

namespace rtb {
namespace {

struct Huff { uint8_t bits[17]; uint8_t vals[256]; int mincode[17], maxcode[18], valptr[17]; bool present = false; };

void build_huff(Huff& h) {
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        h.valptr[l] = k; h.mincode[l] = code;
        code += h.bits[l]; k += h.bits[l];
        h.maxcode[l] = h.bits[l] ? code - 1 : -1;
        code <<= 1;
    }
    h.maxcode[17] = 0x7FFFFFFF;
    h.present = true;
}

struct Reader {
    const uint8_t* p; size_t n, pos = 0;
    uint32_t acc = 0; int cnt = 0;
    bool hit_marker = false; uint8_t marker = 0;
    int bit() {
        if (!cnt) {
            uint8_t b = 0;
            if (!hit_marker && pos < n) {
                b = p[pos++];
                if (b == 0xFF) {
                    uint8_t m = pos < n ? p[pos] : 0xD9;
                    while (m == 0xFF && pos + 1 < n) m = p[++pos];      // fill bytes
                    if (m == 0) ++pos;                                  // stuffed zero
                    else { hit_marker = true; marker = m; ++pos; b = 0; }
                }
            }
            acc = b; cnt = 8;
        }
        --cnt;
        return (acc >> cnt) & 1;
    }
    int bits(int k) { int v = 0; while (k--) v = (v << 1) | bit(); return v; }
    void reset() { cnt = 0; acc = 0; }
};

int decode_symbol(Reader& r, const Huff& h) {
    int code = 0;
    for (int l = 1; l <= 16; ++l) {
        code = (code << 1) | r.bit();
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    throw rt_error(RT_ERR_PARSE, "jpeg: bad Huffman code");
}
int extend(int v, int s) { return s && v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

const uint8_t DEZIGZAG[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                              35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

inline uint8_t clamp8(int x) { return x < 0 ? 0 : (x > 255 ? 255 : uint8_t(x)); }
constexpr int f2f(float x) { return int(double(x * 4096.0f) + 0.5); }       // float constant, scaled, + 0.5, TRUNCATED (so a negative
                                                                            // constant is not the negation of the positive one): as published

// one 1-D pass of the 12-bit fixed-point inverse DCT (the even / odd decomposition of jidctint)
#define RT_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                   \
    int t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;                          \
    p2 = s2; p3 = s6;                                                                \
    p1 = (p2 + p3) * f2f(0.5411961f);                                                 \
    t2 = p1 + p3 * f2f(-1.847759065f);                                              \
    t3 = p1 + p2 * f2f(0.765366865f);                                                 \
    p2 = s0; p3 = s4;                                                                \
    t0 = (p2 + p3) * 4096; t1 = (p2 - p3) * 4096;                                    \
    x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;                          \
    t0 = s7; t1 = s5; t2 = s3; t3 = s1;                                              \
    p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;                          \
    p5 = (p3 + p4) * f2f(1.175875602f);                                               \
    t0 = t0 * f2f(0.298631336f); t1 = t1 * f2f(2.053119869f);                          \
    t2 = t2 * f2f(3.072711026f); t3 = t3 * f2f(1.501321110f);                          \
    p1 = p5 + p1 * f2f(-0.899976223f); p2 = p5 + p2 * f2f(-2.562915447f);          \
    p3 = p3 * f2f(-1.961570560f); p4 = p4 * f2f(-0.390180644f);                    \
    t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;

void idct_block(uint8_t* out, int stride, const short d[64]) {
    int val[64];
    for (int i = 0; i < 8; ++i) {                                        // columns
        const short* c = d + i;
        int* v = val + i;
        if (!c[8] && !c[16] && !c[24] && !c[32] && !c[40] && !c[48] && !c[56]) {
            const int dc = c[0] * 4;
            for (int k = 0; k < 8; ++k) v[8 * k] = dc;
        } else {
            RT_IDCT_1D(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56])
            x0 += 512; x1 += 512; x2 += 512; x3 += 512;
            v[0] = (x0 + t3) >> 10; v[56] = (x0 - t3) >> 10;
            v[8] = (x1 + t2) >> 10; v[48] = (x1 - t2) >> 10;
            v[16] = (x2 + t1) >> 10; v[40] = (x2 - t1) >> 10;
            v[24] = (x3 + t0) >> 10; v[32] = (x3 - t0) >> 10;
        }
    }
    for (int i = 0; i < 8; ++i) {                                        // rows; the level shift (+128) rides on the rounding constant
        const int* v = val + 8 * i;
        uint8_t* o = out + i * stride;
        RT_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        x0 += 65536 + (128 << 17); x1 += 65536 + (128 << 17); x2 += 65536 + (128 << 17); x3 += 65536 + (128 << 17);
        o[0] = clamp8((x0 + t3) >> 17); o[7] = clamp8((x0 - t3) >> 17);
        o[1] = clamp8((x1 + t2) >> 17); o[6] = clamp8((x1 - t2) >> 17);
        o[2] = clamp8((x2 + t1) >> 17); o[5] = clamp8((x2 - t1) >> 17);
        o[3] = clamp8((x3 + t0) >> 17); o[4] = clamp8((x3 - t0) >> 17);
    }
}
#undef RT_IDCT_1D

struct Comp { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, dc_pred = 0, w2 = 0, h2 = 0; std::vector<uint8_t> data; };

// chroma upsampling rows: `near` is the source row nearer to the output row, `far` the other neighbour
void row_v2(uint8_t* out, const uint8_t* near, const uint8_t* far, int w, int) {
    for (int i = 0; i < w; ++i) out[i] = uint8_t((3 * near[i] + far[i] + 2) >> 2);
}
void row_h2(uint8_t* out, const uint8_t* in, const uint8_t*, int w, int) {
    if (w == 1) { out[0] = out[1] = in[0]; return; }
    out[0] = in[0];
    out[1] = uint8_t((in[0] * 3 + in[1] + 2) >> 2);
    int i;
    for (i = 1; i < w - 1; ++i) {
        const int n = 3 * in[i] + 2;
        out[i * 2] = uint8_t((n + in[i - 1]) >> 2);
        out[i * 2 + 1] = uint8_t((n + in[i + 1]) >> 2);
    }
    out[i * 2] = uint8_t((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
    out[i * 2 + 1] = in[w - 1];
}
void row_h2v2(uint8_t* out, const uint8_t* near, const uint8_t* far, int w, int) {
    if (w == 1) { out[0] = out[1] = uint8_t((3 * near[0] + far[0] + 2) >> 2); return; }
    int t1 = 3 * near[0] + far[0];
    out[0] = uint8_t((t1 + 2) >> 2);
    for (int i = 1; i < w; ++i) {
        const int t0 = t1;
        t1 = 3 * near[i] + far[i];
        out[i * 2 - 1] = uint8_t((3 * t0 + t1 + 8) >> 4);
        out[i * 2] = uint8_t((3 * t1 + t0 + 8) >> 4);
    }
    out[w * 2 - 1] = uint8_t((t1 + 2) >> 2);
}
void row_generic(uint8_t* out, const uint8_t* near, const uint8_t*, int w, int hs) {
    for (int i = 0; i < w; ++i) for (int j = 0; j < hs; ++j) out[i * hs + j] = near[i];
}

constexpr int fixed20(float x) { return int(uint32_t(int(x * 4096.0f + 0.5f)) << 8); }

}  // namespace

Bitmap decode_jpeg(const std::string& file, const std::string& what) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(file.data());
    const size_t n = file.size();
    auto fail = [&](int st, const std::string& m) -> rt_error { return rt_error(st, "jpeg " + what + ": " + m); };
    if (n < 4 || p[0] != 0xFF || p[1] != 0xD8) throw fail(RT_ERR_PARSE, "not a JPEG file");
    uint16_t dq[4][64] = {};
    Huff hdc[4], hac[4];
    Comp comp[3];
    int ncomp = 0, width = 0, height = 0, hmax = 1, vmax = 1, restart = 0;
    size_t pos = 2;
    bool have_sof = false;
    for (;;) {
        if (pos + 4 > n) throw fail(RT_ERR_PARSE, "truncated before the scan");
        if (p[pos] != 0xFF) throw fail(RT_ERR_PARSE, "marker expected");
        while (pos < n && p[pos] == 0xFF) ++pos;
        const uint8_t m = p[pos++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) throw fail(RT_ERR_PARSE, "no scan");
        if (pos + 2 > n) throw fail(RT_ERR_PARSE, "truncated segment");
        const size_t len = (size_t(p[pos]) << 8) | p[pos + 1];
        if (len < 2 || pos + len > n) throw fail(RT_ERR_PARSE, "bad segment length");
        const uint8_t* s = p + pos + 2;
        const size_t sl = len - 2;
        if (m == 0xDB) {                                                 // DQT
            size_t i = 0;
            while (i < sl) {
                const int pq = s[i] >> 4, tq = s[i] & 15;
                ++i;
                if (tq > 3 || i + size_t(pq ? 128 : 64) > sl) throw fail(RT_ERR_PARSE, "bad quantisation table");
                for (int k = 0; k < 64; ++k) { dq[tq][DEZIGZAG[k]] = pq ? uint16_t((s[i] << 8) | s[i + 1]) : s[i]; i += pq ? 2 : 1; }
            }
        } else if (m == 0xC4) {                                          // DHT
            size_t i = 0;
            while (i < sl) {
                if (i + 17 > sl) throw fail(RT_ERR_PARSE, "bad Huffman table");
                const int tc = s[i] >> 4, th = s[i] & 15;
                if (tc > 1 || th > 3) throw fail(RT_ERR_PARSE, "bad Huffman table id");
                Huff& h = tc ? hac[th] : hdc[th];
                int total = 0;
                h.bits[0] = 0;
                for (int l = 1; l <= 16; ++l) { h.bits[l] = s[i + l]; total += h.bits[l]; }
                i += 17;
                if (total > 256 || i + size_t(total) > sl) throw fail(RT_ERR_PARSE, "bad Huffman table");
                std::memcpy(h.vals, s + i, size_t(total));
                i += size_t(total);
                build_huff(h);
            }
        } else if (m == 0xC0 || m == 0xC1) {                             // SOF0 / SOF1
            if (sl < 6) throw fail(RT_ERR_PARSE, "bad frame header");
            if (s[0] != 8) throw fail(RT_ERR_UNSUPPORTED, "only 8-bit samples");
            height = (s[1] << 8) | s[2]; width = (s[3] << 8) | s[4]; ncomp = s[5];
            if (!width || !height) throw fail(RT_ERR_PARSE, "empty image");
            if (ncomp != 1 && ncomp != 3) throw fail(RT_ERR_UNSUPPORTED, "only 1 or 3 components");
            if (sl < size_t(6 + 3 * ncomp)) throw fail(RT_ERR_PARSE, "bad frame header");
            for (int c = 0; c < ncomp; ++c) {
                comp[c].id = s[6 + 3 * c]; comp[c].h = s[7 + 3 * c] >> 4; comp[c].v = s[7 + 3 * c] & 15; comp[c].tq = s[8 + 3 * c];
                if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4 || comp[c].tq > 3) throw fail(RT_ERR_PARSE, "bad component");
                hmax = std::max(hmax, comp[c].h); vmax = std::max(vmax, comp[c].v);
            }
            have_sof = true;
        } else if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
            throw fail(RT_ERR_UNSUPPORTED, "only baseline / extended sequential Huffman JPEG (no progressive, lossless or arithmetic coding)");
        } else if (m == 0xDD) {
            if (sl < 2) throw fail(RT_ERR_PARSE, "bad restart interval");
            restart = (s[0] << 8) | s[1];
        } else if (m == 0xDA) {                                          // SOS
            if (!have_sof) throw fail(RT_ERR_PARSE, "scan before frame header");
            if (sl < 1 || s[0] != ncomp || sl < size_t(1 + 2 * ncomp + 3)) throw fail(RT_ERR_UNSUPPORTED, "only one interleaved scan");
            for (int k = 0; k < ncomp; ++k) {
                int c = 0;
                while (c < ncomp && comp[c].id != s[1 + 2 * k]) ++c;
                if (c == ncomp) throw fail(RT_ERR_PARSE, "scan names an unknown component");
                comp[c].td = s[2 + 2 * k] >> 4; comp[c].ta = s[2 + 2 * k] & 15;
                if (comp[c].td > 3 || comp[c].ta > 3 || !hdc[comp[c].td].present || !hac[comp[c].ta].present) throw fail(RT_ERR_PARSE, "scan names a missing Huffman table");
            }
            pos += len;
            break;
        }
        pos += len;
    }
    for (int c = 0; c < ncomp; ++c)
        if (hmax % comp[c].h || vmax % comp[c].v) throw fail(RT_ERR_UNSUPPORTED, "fractional sampling ratios");

    // ---- entropy-coded data: interleaved MCUs ----
    const int mcu_w = 8 * hmax, mcu_h = 8 * vmax;
    const int mcus_x = (width + mcu_w - 1) / mcu_w, mcus_y = (height + mcu_h - 1) / mcu_h;
    for (int c = 0; c < ncomp; ++c) {
        comp[c].w2 = mcus_x * comp[c].h * 8; comp[c].h2 = mcus_y * comp[c].v * 8;
        comp[c].data.assign(size_t(comp[c].w2) * comp[c].h2, 0);
    }
    Reader r{p, n, pos};
    int todo = restart ? restart : 0x7FFFFFFF;
    for (int my = 0; my < mcus_y; ++my)
        for (int mx = 0; mx < mcus_x; ++mx) {
            for (int c = 0; c < ncomp; ++c)
                for (int by = 0; by < comp[c].v; ++by)
                    for (int bx = 0; bx < comp[c].h; ++bx) {
                        short d[64] = {};
                        const int t = decode_symbol(r, hdc[comp[c].td]);
                        if (t > 15) throw fail(RT_ERR_PARSE, "bad DC category");
                        const int diff = t ? extend(r.bits(t), t) : 0;
                        comp[c].dc_pred += diff;
                        d[0] = short(comp[c].dc_pred * dq[comp[c].tq][0]);
                        for (int k = 1; k < 64;) {
                            const int rs = decode_symbol(r, hac[comp[c].ta]);
                            const int sz = rs & 15, run = rs >> 4;
                            if (!sz) { if (rs != 0xF0) break; k += 16; continue; }
                            k += run;
                            if (k > 63) throw fail(RT_ERR_PARSE, "coefficient index out of range");
                            const int z = DEZIGZAG[k++];
                            d[z] = short(extend(r.bits(sz), sz) * dq[comp[c].tq][z]);
                        }
                        const int x0 = (mx * comp[c].h + bx) * 8, y0 = (my * comp[c].v + by) * 8;
                        idct_block(comp[c].data.data() + size_t(y0) * comp[c].w2 + x0, comp[c].w2, d);
                    }
            if (--todo <= 0) {                                            // restart interval: byte-align, expect RSTn, reset predictors
                r.reset();
                if (!r.hit_marker) { (void)r.bit(); r.reset(); }          // pull the marker in
                if (!(r.hit_marker && r.marker >= 0xD0 && r.marker <= 0xD7)) {
                    if (my == mcus_y - 1 && mx == mcus_x - 1) break;
                    throw fail(RT_ERR_PARSE, "restart marker expected");
                }
                r.hit_marker = false;
                for (int c = 0; c < ncomp; ++c) comp[c].dc_pred = 0;
                todo = restart;
            }
        }

    // ---- upsample and convert ----
    Bitmap bm;
    bm.w = uint32_t(width); bm.h = uint32_t(height);
    bm.rgb.resize(size_t(width) * height * 3);
    std::vector<uint8_t> line[3];
    for (int c = 0; c < ncomp; ++c) line[c].resize(size_t(width) + 16 * 4);
    for (int y = 0; y < height; ++y) {
        const uint8_t* src[3] = {nullptr, nullptr, nullptr};
        for (int c = 0; c < ncomp; ++c) {
            const int hs = hmax / comp[c].h, vs = vmax / comp[c].v;
            const int w_lores = (width + hs - 1) / hs;
            // source rows: for vs == 2 the output row y lies between source rows; `near` is the closer one
            int ynear = y / vs, yfar = ynear;
            const int rows = (height * comp[c].v + vmax - 1) / vmax;      // the component's own height
            if (vs == 2) yfar = (y & 1) ? std::min(ynear + 1, rows - 1) : std::max(ynear - 1, 0);
            const uint8_t* near = comp[c].data.data() + size_t(ynear) * comp[c].w2;
            const uint8_t* far = comp[c].data.data() + size_t(yfar) * comp[c].w2;
            if (hs == 1 && vs == 1) { src[c] = near; continue; }
            auto fn = (hs == 1 && vs == 2) ? row_v2 : (hs == 2 && vs == 1) ? row_h2 : (hs == 2 && vs == 2) ? row_h2v2 : row_generic;
            if (fn == row_generic && vs != 1) near = comp[c].data.data() + size_t(y / vs) * comp[c].w2;
            fn(line[c].data(), near, far, w_lores, hs);
            src[c] = line[c].data();
        }
        uint8_t* out = bm.rgb.data() + size_t(y) * width * 3;
        if (ncomp == 1) {
            for (int x = 0; x < width; ++x) { out[3 * x] = out[3 * x + 1] = out[3 * x + 2] = src[0][x]; }
        } else {
            for (int x = 0; x < width; ++x) {                             // 20-bit fixed point, rounding constant on the luma term
                const int yf = (int(src[0][x]) << 20) + (1 << 19);
                const int cb = int(src[1][x]) - 128, cr = int(src[2][x]) - 128;
                int rr = yf + cr * fixed20(1.40200f);
                int gg = yf + cr * -fixed20(0.71414f) + int(uint32_t(cb * -fixed20(0.34414f)) & 0xFFFF0000u);
                int bb = yf + cb * fixed20(1.77200f);
                rr >>= 20; gg >>= 20; bb >>= 20;
                out[3 * x] = clamp8(rr); out[3 * x + 1] = clamp8(gg); out[3 * x + 2] = clamp8(bb);
            }
        }
    }
    return bm;
}

}  // namespace rtb
