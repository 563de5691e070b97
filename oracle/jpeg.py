"""ORACLE (test infrastructure): baseline JPEG decoding restated from the published algorithm of the decoder the
reference uses.

The reference loads every frame with ``PIL.Image.open(...)`` (dataset/local_dataset.py:283-286). Pillow hands the file to
the IJG / libjpeg-turbo decoder (third-party, not under /root/reference; Pillow 12.x in this image bundles libjpeg-turbo
3.x) with its default settings: integer "islow" inverse DCT (jidctint.c), "fancy" triangle-filter chroma upsampling
(jdsample.c h2v1 / h2v2_fancy_upsample) and the 16-bit fixed-point YCbCr -> RGB conversion (jdcolor.c). All three are
exact integer algorithms, so a restatement can be - and is required to be - bit-exact:

    decode(bytes) == numpy.asarray(PIL.Image.open(BytesIO(bytes)).convert("RGB"))

``tests/test_jpeg.py`` pins this module against Pillow itself (CPU, many sizes / qualities / subsamplings / restart
intervals) and then uses it as the checker of the product path (host Huffman decoding in csrc/jpeg_host.cu + GPU
dequantisation / IDCT / upsampling / colour kernels in csrc/jpeg.cu).

Supported: baseline sequential DCT (SOF0), 8-bit samples, 1 or 3 components, sampling factors 1x1 / 2x1 / 2x2 for
luma with 1x1 chroma (4:4:4, 4:2:2, 4:2:0), restart intervals, JFIF YCbCr. Anything else raises ValueError, like the
product (ffmpeg / OpenCV frame dumps such as the reference's datasets are baseline 4:2:0).
"""
import numpy as np

ZIGZAG = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
    62, 63], dtype=np.int64)      # natural index of the k-th zigzag coefficient


class _Bits:
    """MSB-first bit reader over the entropy-coded segment with 0xFF00 unstuffing; stops at markers."""

    def __init__(self, data, pos):
        self.d, self.pos, self.acc, self.n = data, pos, 0, 0

    def _fill(self):
        while self.n <= 24:
            if self.pos >= len(self.d):
                b = 0
            else:
                b = self.d[self.pos]
                if b == 0xFF:
                    nxt = self.d[self.pos + 1] if self.pos + 1 < len(self.d) else 0
                    if nxt == 0:
                        self.pos += 2
                    else:
                        b = 0            # a marker: feed zeros, do not advance
                else:
                    self.pos += 1
            self.acc = ((self.acc << 8) | b) & 0xFFFFFFFFFF
            self.n += 8

    def get(self, k):
        if k == 0:
            return 0
        if self.n < k:
            self._fill()
        self.n -= k
        return (self.acc >> self.n) & ((1 << k) - 1)

    def align_and_skip_restart(self):
        self.acc = self.n = 0
        # the byte position is at the marker (fill never consumes one)
        while self.pos + 1 < len(self.d) and not (self.d[self.pos] == 0xFF and 0xD0 <= self.d[self.pos + 1] <= 0xD7):
            self.pos += 1
        self.pos += 2


def _huff_table(counts, symbols):
    """(maxcode[17], valptr[17], mincode[17], symbols) of the canonical code (ITU T.81 Annex C / F.2.2.3)."""
    codes, code, k = [], 0, 0
    mincode, maxcode, valptr = [0] * 17, [-1] * 18, [0] * 17
    for length in range(1, 17):
        valptr[length] = k
        mincode[length] = code
        for _ in range(counts[length - 1]):
            codes.append(code)
            code += 1
            k += 1
        maxcode[length] = code - 1 if counts[length - 1] else -1
        code <<= 1
    return maxcode, valptr, mincode, symbols


def _decode_symbol(bits, tab):
    maxcode, valptr, mincode, symbols = tab
    code = 0
    for length in range(1, 17):
        code = (code << 1) | bits.get(1)
        if maxcode[length] >= 0 and code <= maxcode[length] and code >= mincode[length]:
            return symbols[valptr[length] + code - mincode[length]]
    raise ValueError("bad Huffman code")


def _extend(v, t):
    return v if v >= (1 << (t - 1)) else v - (1 << t) + 1


def parse(data):
    """Markers -> dict(width, height, comps=[(id, h, v, tq, td, ta)], qt, dc, ac, restart, scan_pos)."""
    d = data if isinstance(data, (bytes, bytearray)) else bytes(data)
    if d[0:2] != b"\xff\xd8":
        raise ValueError("not a JPEG (no SOI)")
    pos, qt, dc, ac, restart = 2, {}, {}, {}, 0
    frame = None
    while True:
        if d[pos] != 0xFF:
            raise ValueError("marker expected")
        while d[pos + 1] == 0xFF:
            pos += 1
        m = d[pos + 1]
        pos += 2
        if m == 0xD9:
            raise ValueError("EOI before SOS")
        length = (d[pos] << 8) | d[pos + 1]
        seg = d[pos + 2:pos + length]
        if m == 0xDB:
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                if pq:
                    raise ValueError("16-bit quantisation tables are not baseline")
                t = np.zeros(64, np.int32)
                t[ZIGZAG] = np.frombuffer(seg[i + 1:i + 65], np.uint8)
                qt[tq] = t
                i += 65
        elif m == 0xC0:
            if seg[0] != 8:
                raise ValueError("only 8-bit samples")
            h, w, nc = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = [[seg[6 + 3 * i], seg[7 + 3 * i] >> 4, seg[7 + 3 * i] & 15, seg[8 + 3 * i], 0, 0] for i in range(nc)]
            frame = (w, h, comps)
        elif m in (0xC1, 0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise ValueError("only baseline sequential JPEG (SOF0) is supported")
        elif m == 0xC4:
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                counts = list(seg[i + 1:i + 17])
                n = sum(counts)
                (ac if tc else dc)[th] = _huff_table(counts, list(seg[i + 17:i + 17 + n]))
                i += 17 + n
        elif m == 0xDD:
            restart = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            if frame is None:
                raise ValueError("SOS before SOF")
            ns = seg[0]
            w, h, comps = frame
            if ns != len(comps):
                raise ValueError("non-interleaved scans are not supported")
            for i in range(ns):
                cid, tt = seg[1 + 2 * i], seg[2 + 2 * i]
                for c in comps:
                    if c[0] == cid:
                        c[4], c[5] = tt >> 4, tt & 15
            return dict(width=w, height=h, comps=[tuple(c) for c in comps], qt=qt, dc=dc, ac=ac, restart=restart,
                        scan_pos=pos + length)
        pos += length


def decode_coefficients(data):
    """Entropy decoding: per component an int16 array [blocks_v][blocks_h][64] of quantised coefficients in natural
    order (padded to whole MCUs) plus the header dict."""
    hd = parse(data)
    comps = hd["comps"]
    if len(comps) not in (1, 3):
        raise ValueError("1 or 3 components")
    hmax, vmax = max(c[1] for c in comps), max(c[2] for c in comps)
    if len(comps) == 3 and (comps[1][1:3] != (1, 1) or comps[2][1:3] != (1, 1) or (hmax, vmax) not in ((1, 1), (2, 1), (2, 2))):
        raise ValueError("unsupported sampling factors")
    if len(comps) == 1:
        hmax = vmax = 1
        comps = [(comps[0][0], 1, 1) + comps[0][3:]]
        hd["comps"] = comps
    mcux, mcuy = -(-hd["width"] // (8 * hmax)), -(-hd["height"] // (8 * vmax))
    coefs = [np.zeros((mcuy * c[2], mcux * c[1], 64), np.int16) for c in comps]
    bits = _Bits(data if isinstance(data, (bytes, bytearray)) else bytes(data), hd["scan_pos"])
    pred = [0] * len(comps)
    count = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if hd["restart"] and count and count % hd["restart"] == 0:
                bits.align_and_skip_restart()
                pred = [0] * len(comps)
            count += 1
            for ci, c in enumerate(comps):
                for by in range(c[2]):
                    for bx in range(c[1]):
                        blk = coefs[ci][my * c[2] + by, mx * c[1] + bx]
                        t = _decode_symbol(bits, hd["dc"][c[4]])
                        diff = _extend(bits.get(t), t) if t else 0
                        pred[ci] += diff
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = _decode_symbol(bits, hd["ac"][c[5]])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = _extend(bits.get(s), s)
                            k += 1
    hd.update(hmax=hmax, vmax=vmax, mcux=mcux, mcuy=mcuy)
    return coefs, hd


# ---- jidctint.c (jpeg_idct_islow): CONST_BITS = 13, PASS1_BITS = 2
_F = dict(c0_298631336=2446, c0_390180644=3196, c0_541196100=4433, c0_765366865=6270, c0_899976223=7373,
          c1_175875602=9633, c1_501321110=12299, c1_847759065=15137, c1_961570560=16069, c2_053119869=16819,
          c2_562915447=20995, c3_072711026=25172)


def _idct_1d(v, shift_out, pass1):
    """One pass over the LAST axis of an int64 array [..., 8]; returns the descaled int64 outputs."""
    z2, z3 = v[..., 2], v[..., 6]
    z1 = (z2 + z3) * _F["c0_541196100"]
    tmp2 = z1 + z3 * (-_F["c1_847759065"])
    tmp3 = z1 + z2 * _F["c0_765366865"]
    z2, z3 = v[..., 0], v[..., 4]
    tmp0 = (z2 + z3) << 13
    tmp1 = (z2 - z3) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = v[..., 7], v[..., 5], v[..., 3], v[..., 1]
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * _F["c1_175875602"]
    t0 = t0 * _F["c0_298631336"]
    t1 = t1 * _F["c2_053119869"]
    t2 = t2 * _F["c3_072711026"]
    t3 = t3 * _F["c1_501321110"]
    z1 = z1 * (-_F["c0_899976223"])
    z2 = z2 * (-_F["c2_562915447"])
    z3 = z3 * (-_F["c1_961570560"]) + z5
    z4 = z4 * (-_F["c0_390180644"]) + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    rnd = 1 << (shift_out - 1)
    out = np.stack([tmp10 + t3, tmp11 + t2, tmp12 + t1, tmp13 + t0, tmp13 - t0, tmp12 - t1, tmp11 - t2, tmp10 - t3], -1)
    return (out + rnd) >> shift_out


def idct_blocks(coef, qt):
    """int16 [..., 64] quantised coefficients -> uint8 [..., 8, 8] samples (dequantise, islow IDCT, +128, range limit)."""
    w = coef.astype(np.int64).reshape(coef.shape[:-1] + (8, 8)) * qt.astype(np.int64).reshape(8, 8)
    ws = _idct_1d(np.swapaxes(w, -1, -2), 13 - 2, True)            # pass 1: columns, results scaled up by 2^PASS1_BITS
    ws = np.swapaxes(ws, -1, -2)
    out = _idct_1d(ws, 13 + 2 + 3, False)                          # pass 2: rows
    idx = out & 1023                                               # RANGE_MASK, then the wrap-around range-limit table
    s = np.where(idx < 512, idx, idx - 1024)
    return np.clip(s + 128, 0, 255).astype(np.uint8)


def _plane(samples, real_h, real_w):
    """[bv][bh][8][8] -> [bv*8][bh*8] cropped to the component's real size."""
    bv, bh = samples.shape[:2]
    return samples.transpose(0, 2, 1, 3).reshape(bv * 8, bh * 8)[:real_h, :real_w]


def _h2v1_fancy(p):
    """jdsample.c h2v1_fancy_upsample on every row: [h][w] -> [h][2w]."""
    x = p.astype(np.int32)
    h, w = x.shape
    out = np.empty((h, 2 * w), np.int32)
    if w == 1:
        out[:, 0] = out[:, 1] = x[:, 0]
        return out.astype(np.uint8)
    left = np.concatenate([x[:, :1], x[:, :-1]], 1)
    right = np.concatenate([x[:, 1:], x[:, -1:]], 1)
    out[:, 0::2] = (x * 3 + left + 1) >> 2
    out[:, 1::2] = (x * 3 + right + 2) >> 2
    out[:, 0] = x[:, 0]
    out[:, -1] = x[:, -1]
    return out.astype(np.uint8)


def _h2v2_fancy(p):
    """jdsample.c h2v2_fancy_upsample with jdmainct.c's edge rows (first / last real row replicated): [h][w] -> [2h][2w]."""
    x = p.astype(np.int32)
    h, w = x.shape
    above = np.concatenate([x[:1], x[:-1]], 0)
    below = np.concatenate([x[1:], x[-1:]], 0)
    out = np.empty((2 * h, 2 * w), np.int32)
    for v, other in ((0, above), (1, below)):
        cs = x * 3 + other                                      # column sums (vertical 3/4 + 1/4)
        if w == 1:
            row = np.concatenate([(cs * 4 + 8) >> 4, (cs * 4 + 7) >> 4], 1)
        else:
            last = np.concatenate([cs[:, :1], cs[:, :-1]], 1)
            nxt = np.concatenate([cs[:, 1:], cs[:, -1:]], 1)
            row = np.empty((h, 2 * w), np.int32)
            row[:, 0::2] = (cs * 3 + last + 8) >> 4
            row[:, 1::2] = (cs * 3 + nxt + 7) >> 4
            row[:, 0] = (cs[:, 0] * 4 + 8) >> 4
            row[:, -1] = (cs[:, -1] * 4 + 7) >> 4
        out[v::2] = row
    return out.astype(np.uint8)


def _fix(x):
    return int(x * 65536 + 0.5)


_X = np.arange(256, dtype=np.int64) - 128
CR_R = (_fix(1.40200) * _X + 32768) >> 16
CB_B = (_fix(1.77200) * _X + 32768) >> 16
CR_G = -_fix(0.71414) * _X
CB_G = -_fix(0.34414) * _X + 32768


def ycc_to_rgb(y, cb, cr):
    """jdcolor.c ycc_rgb_convert on uint8 planes."""
    yy = y.astype(np.int64)
    r = yy + CR_R[cr]
    g = yy + ((CB_G[cb] + CR_G[cr]) >> 16)
    b = yy + CB_B[cb]
    return np.clip(np.stack([r, g, b], -1), 0, 255).astype(np.uint8)


def decode(data):
    """JPEG bytes -> uint8 [H][W][3] RGB (grayscale files are replicated to 3 channels like ``convert('RGB')``)."""
    coefs, hd = decode_coefficients(data)
    W, H, comps = hd["width"], hd["height"], hd["comps"]
    planes = []
    for ci, c in enumerate(comps):
        cw, ch = -(-W * c[1] // hd["hmax"]), -(-H * c[2] // hd["vmax"])
        planes.append(_plane(idct_blocks(coefs[ci], hd["qt"][c[3]]), ch, cw))
    if len(comps) == 1:
        return np.repeat(planes[0][:, :, None], 3, 2)
    y, cb, cr = planes
    if (hd["hmax"], hd["vmax"]) == (2, 2):
        cb, cr = _h2v2_fancy(cb), _h2v2_fancy(cr)
    elif (hd["hmax"], hd["vmax"]) == (2, 1):
        cb, cr = _h2v1_fancy(cb), _h2v1_fancy(cr)
    return ycc_to_rgb(y, cb[:H, :W], cr[:H, :W])
