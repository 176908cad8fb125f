"""TEST INFRASTRUCTURE: an independent (pure Python) restatement of the baseline-JPEG arithmetic the reference's bitmap loader
goes through (scene/texture/bitmap.hpp:15 -> stb_image: 12-bit fixed-point inverse DCT, 20-bit fixed-point YCbCr -> RGB), for
4:4:4 / grey-scale baseline files without restart markers - which is what scenes/hw12/textures/dragon.jpg is.  The fixtures'
texel bytes come from here; the product's C++ decoder (simd-raytracer_b200/host/jpeg_decode.cpp) is checked against it byte for
byte, and both against the bitmap quadrant of the reference's published outputs/textures.png (tests/test_oracle_golden.py)."""
from __future__ import annotations

import numpy as np

DEZIGZAG = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
            35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]


def _f2f(x: float) -> int:
    """(int)(float(x) * 4096 + 0.5): float constant, truncated toward zero (a negative constant is not the negated positive one)"""
    return int(float(np.float32(x)) * 4096 + 0.5)


C = {k: _f2f(v) for k, v in dict(a=0.5411961, b=-1.847759065, c=0.765366865, d=1.175875602, e=0.298631336, f=2.053119869, g=3.072711026,
                                 h=1.501321110, i=-0.899976223, j=-2.562915447, k=-1.961570560, l=-0.390180644).items()}


def _idct_1d(s0, s1, s2, s3, s4, s5, s6, s7):
    p2, p3 = s2, s6
    p1 = (p2 + p3) * C["a"]
    t2 = p1 + p3 * C["b"]
    t3 = p1 + p2 * C["c"]
    p2, p3 = s0, s4
    t0, t1 = (p2 + p3) * 4096, (p2 - p3) * 4096
    x0, x3, x1, x2 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    t0, t1, t2, t3 = s7, s5, s3, s1
    p3, p4, p1, p2 = t0 + t2, t1 + t3, t0 + t3, t1 + t2
    p5 = (p3 + p4) * C["d"]
    t0, t1, t2, t3 = t0 * C["e"], t1 * C["f"], t2 * C["g"], t3 * C["h"]
    p1, p2 = p5 + p1 * C["i"], p5 + p2 * C["j"]
    p3, p4 = p3 * C["k"], p4 * C["l"]
    return x0, x1, x2, x3, t0 + p1 + p3, t1 + p2 + p4, t2 + p2 + p3, t3 + p1 + p4


def _clamp(x: int) -> int:
    return 0 if x < 0 else (255 if x > 255 else x)


def _idct_block(d):
    v = [0] * 64
    for i in range(8):
        c = d[i::8]
        if not any(c[1:]):
            for k in range(8):
                v[8 * k + i] = c[0] * 4
        else:
            x0, x1, x2, x3, t0, t1, t2, t3 = _idct_1d(*c)
            x0 += 512; x1 += 512; x2 += 512; x3 += 512
            for k, val in ((0, x0 + t3), (7, x0 - t3), (1, x1 + t2), (6, x1 - t2), (2, x2 + t1), (5, x2 - t1), (3, x3 + t0), (4, x3 - t0)):
                v[8 * k + i] = val >> 10
    out = [0] * 64
    r = 65536 + (128 << 17)
    for i in range(8):
        x0, x1, x2, x3, t0, t1, t2, t3 = _idct_1d(*v[8 * i:8 * i + 8])
        x0 += r; x1 += r; x2 += r; x3 += r
        for k, val in ((0, x0 + t3), (7, x0 - t3), (1, x1 + t2), (6, x1 - t2), (2, x2 + t1), (5, x2 - t1), (3, x3 + t0), (4, x3 - t0)):
            out[8 * i + k] = _clamp(val >> 17)
    return out


class _Bits:
    def __init__(self, data: bytes, pos: int):
        self.d, self.pos, self.acc, self.cnt = data, pos, 0, 0

    def bit(self) -> int:
        if not self.cnt:
            b = self.d[self.pos] if self.pos < len(self.d) else 0
            self.pos += 1
            if b == 0xFF:
                if self.pos < len(self.d) and self.d[self.pos] == 0:
                    self.pos += 1
                else:
                    b = 0                   # a marker: the scan is over, feed zeros
                    self.pos = len(self.d)
            self.acc, self.cnt = b, 8
        self.cnt -= 1
        return (self.acc >> self.cnt) & 1

    def bits(self, n: int) -> int:
        v = 0
        for _ in range(n):
            v = (v << 1) | self.bit()
        return v


def _huff(bits, vals):
    table, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            table[(length, code)] = vals[k]
            code += 1
            k += 1
        code <<= 1
    return table


def _symbol(r: _Bits, table) -> int:
    code = 0
    for length in range(1, 17):
        code = (code << 1) | r.bit()
        if (length, code) in table:
            return table[(length, code)]
    raise ValueError("bad Huffman code")


def _extend(v: int, s: int) -> int:
    return v - (1 << s) + 1 if s and v < (1 << (s - 1)) else v


def decode(data: bytes) -> np.ndarray:
    """(h, w, 3) uint8; raises NotImplementedError for anything but 8-bit baseline 4:4:4 / grey-scale without restart markers"""
    if data[:2] != b"\xff\xd8":
        raise ValueError("not a JPEG")
    dq, hdc, hac, comps, pos = {}, {}, {}, [], 2
    width = height = 0
    while True:
        if data[pos] != 0xFF:
            raise ValueError("marker expected")
        m = data[pos + 1]
        length = (data[pos + 2] << 8) | data[pos + 3]
        seg = data[pos + 4:pos + 2 + length]
        if m == 0xDB:
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                if pq:
                    raise NotImplementedError("16-bit quantisation table")
                t = [0] * 64
                for k in range(64):
                    t[DEZIGZAG[k]] = seg[i + 1 + k]
                dq[tq] = t
                i += 65
        elif m == 0xC4:
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                bits = list(seg[i + 1:i + 17])
                n = sum(bits)
                (hac if tc else hdc)[th] = _huff(bits, list(seg[i + 17:i + 17 + n]))
                i += 17 + n
        elif m == 0xC0:
            if seg[0] != 8:
                raise NotImplementedError("sample precision")
            height, width, nc = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            for c in range(nc):
                if seg[7 + 3 * c] != 0x11:
                    raise NotImplementedError("subsampled chroma")
                comps.append({"id": seg[6 + 3 * c], "tq": seg[8 + 3 * c], "pred": 0})
            if nc not in (1, 3):
                raise NotImplementedError("component count")
        elif m in (0xC1, 0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF, 0xDD):
            raise NotImplementedError(f"marker {m:#x}")
        elif m == 0xDA:
            for k in range(seg[0]):
                c = next(c for c in comps if c["id"] == seg[1 + 2 * k])
                c["td"], c["ta"] = seg[2 + 2 * k] >> 4, seg[2 + 2 * k] & 15
            pos += 2 + length
            break
        pos += 2 + length
    bw, bh = (width + 7) // 8, (height + 7) // 8
    planes = [np.zeros((bh * 8, bw * 8), np.uint8) for _ in comps]
    r = _Bits(data, pos)
    for by in range(bh):
        for bx in range(bw):
            for ci, c in enumerate(comps):
                d = [0] * 64
                t = _symbol(r, hdc[c["td"]])
                c["pred"] += _extend(r.bits(t), t) if t else 0
                q = dq[c["tq"]]
                d[0] = c["pred"] * q[0]
                k = 1
                while k < 64:
                    rs = _symbol(r, hac[c["ta"]])
                    s, run = rs & 15, rs >> 4
                    if not s:
                        if rs != 0xF0:
                            break
                        k += 16
                        continue
                    k += run
                    z = DEZIGZAG[k]
                    k += 1
                    d[z] = _extend(r.bits(s), s) * q[z]
                planes[ci][by * 8:by * 8 + 8, bx * 8:bx * 8 + 8] = np.array(_idct_block(d), np.uint8).reshape(8, 8)
    if len(comps) == 1:
        g = planes[0][:height, :width]
        return np.stack([g, g, g], axis=2)

    def fx(x):
        return int(np.float32(x) * np.float32(4096.0) + np.float32(0.5)) << 8

    y = planes[0][:height, :width].astype(np.int64)
    cb = planes[1][:height, :width].astype(np.int64) - 128
    cr = planes[2][:height, :width].astype(np.int64) - 128
    yf = (y << 20) + (1 << 19)

    def i32(a):                          # wrap to a 32-bit two's-complement int, as the C expression does
        return ((a + (1 << 31)) % (1 << 32)) - (1 << 31)

    rr = yf + cr * fx(1.40200)
    gg = yf + cr * -fx(0.71414) + i32((i32(cb * -fx(0.34414)) % (1 << 32)) & 0xFFFF0000)
    bb = yf + cb * fx(1.77200)
    out = np.stack([rr >> 20, gg >> 20, bb >> 20], axis=2)
    return np.clip(out, 0, 255).astype(np.uint8)
