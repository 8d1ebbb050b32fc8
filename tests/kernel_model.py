"""Lane-level numpy model of csrc/conv_stage.cu (test infrastructure).

It mirrors the kernel's work decomposition literally -- warp tasks, lane -> region mapping, integer
pooled sums, quad shuffles, conv3 reduce-scatter, feature offsets, hi/lo fp16 split -- so that the index
arithmetic of the CUDA code can be checked against the oracle on a box without a GPU.  Any change to the
mapping in conv_stage.cu must be mirrored here."""
import numpy as np

kOffC3S, kOffC3M, kOffC3L, kOffC2S, kOffC2M, kOffC2L = 0, 512, 640, 672, 2208, 2592
kW1Off, kB1Off, kW2Off, kB2Off, kW3Off, kW3Stride, kB3Off, kBr = 0, 256, 272, 1808, 1832, 776, 4936, 4968
F = np.float32


def leaky(v):
    return np.maximum(F(0.2) * v, v)


def lane_program(tile, oy, ox, P, wb, cst, d):
    """One lane: tile = [64,64] uint8 CTU, region origin (oy, ox) in pixels, pool P, branch weights wb
    (flat float32 block), quad position d.  Returns (region_sum, closure computing c2/part given wsum)."""
    reg = tile[oy:oy + 8 * P, ox:ox + 8 * P].astype(np.int64)
    rsum = int(reg.sum())

    def rest(wsum):
        acc2 = wb[kB2Off:kB2Off + 24].copy()
        for patch in range(4):
            py, px = patch >> 1, patch & 1
            p0 = reg[4 * P * py:4 * P * py + 4 * P, 4 * P * px:4 * P * px + 4 * P]
            ps = p0.reshape(4, P, 4, P).sum(axis=(1, 3)).reshape(16)          # s[ky*4+kx]
            a1 = wb[kB1Off:kB1Off + 16].copy()
            for t in range(16):
                x = F(F(int(ps[t]) * 256 - wsum) * cst)
                a1 = (x * wb[kW1Off + t * 16:kW1Off + t * 16 + 16] + a1).astype(F)
            for ci in range(16):
                c = leaky(a1[ci])
                o = kW2Off + (patch * 16 + ci) * 24
                acc2 = (c * wb[o:o + 24] + acc2).astype(F)
        c2 = leaky(acc2)
        part = np.zeros(32, F)
        for og in range(4):                                   # rolled over the four 8-channel output groups
            for ci in range(24):
                o = kW3Off + d * kW3Stride + (og * 24 + ci) * 8
                part[8 * og:8 * og + 8] = (c2[ci] * wb[o:o + 8] + part[8 * og:8 * og + 8]).astype(F)
        return c2, part
    return rsum, rest


def conv_features_group(tiles, convw, cst3):
    """tiles: [16,64,64] uint8 (one tile group).  convw: [3,4968] packed conv weights (S, M, L).
    Returns float32 [16, 2688] features computed task by task exactly like the kernel: 21 warp tasks,
    every lane owning two regions (A, B)."""
    feat = np.full((16, 2688), np.nan, F)
    for task in range(21):
        regs = []   # per lane: [(c, ry, rx, c2_off, c3_off) for A, B], plus d, P, br
        for lane in range(32):
            d = lane & 3
            if task < 16:
                q = lane >> 2
                qy, qx = q >> 2, q & 3
                ry_a, rx = 2 * qy + (d >> 1), 2 * qx + (d & 1)
                ry_b = ry_a + 4
                P, br = 1, 0
                A = (task, ry_a, rx, kOffC2S + (ry_a * 8 + rx) * 24, kOffC3S + (qy * 4 + qx) * 32)
                B = (task, ry_b, rx, kOffC2S + (ry_b * 8 + rx) * 24, kOffC3S + ((qy + 2) * 4 + qx) * 32)
            elif task < 20:
                l16 = lane & 15
                q = l16 >> 2
                qy, qx = q >> 1, q & 1
                ry, rx = 2 * qy + (d >> 1), 2 * qx + (d & 1)
                ca = 4 * (task - 16) + (lane >> 4)
                P, br = 2, 1
                A = (ca, ry, rx, kOffC2M + (ry * 4 + rx) * 24, kOffC3M + (qy * 2 + qx) * 32)
                B = (ca + 2,) + A[1:]
            else:
                ca = lane >> 2
                ry, rx = d >> 1, d & 1
                P, br = 4, 2
                A = (ca, ry, rx, kOffC2L + (ry * 2 + rx) * 24, kOffC3L)
                B = (ca + 8,) + A[1:]
            regs.append((d, P, br, A, B))
        for which in (3, 4):     # region A of every lane, then region B: the quad exchanges stay inside one of them
            lanes = []
            for lane in range(32):
                d, P, br = regs[lane][:3]
                c, ry, rx, c2_off, c3_off = regs[lane][which]
                rsum, rest = lane_program(tiles[c], 8 * P * ry, 8 * P * rx, P, convw[br], cst3[br], d)
                lanes.append(dict(c=c, d=d, rsum=rsum, rest=rest, c2_off=c2_off, c3_off=c3_off, wb=convw[br]))
            r1 = [lanes[l]["rsum"] + lanes[l ^ 1]["rsum"] for l in range(32)]
            wsum = [r1[l] + r1[l ^ 2] for l in range(32)]
            parts = []
            for l in range(32):
                c2, part = lanes[l]["rest"](wsum[l])
                feat[lanes[l]["c"], lanes[l]["c2_off"]:lanes[l]["c2_off"] + 24] = c2
                parts.append(part)
            # per output group: all-reduce over the quad (xor 1, then xor 2); lane d keeps group d
            for l in range(32):
                d = lanes[l]["d"]
                g = slice(8 * d, 8 * d + 8)
                s1 = {m: (parts[m][g] + parts[m ^ 1][g]).astype(F) for m in (l, l ^ 2)}
                r8 = (s1[l] + s1[l ^ 2]).astype(F)
                b3 = lanes[l]["wb"][kB3Off + 8 * d:kB3Off + 8 * d + 8]
                o = lanes[l]["c3_off"] + 8 * d
                feat[lanes[l]["c"], o:o + 8] = leaky((r8 + b3).astype(F))
    return feat


def split_hi_lo(v, exp):
    s = (v.astype(F) * F(2.0 ** exp)).astype(F)
    hi = s.astype(np.float16)
    lo = (s - hi.astype(F)).astype(np.float16)
    return hi, lo


def fc1_three_pass(feat, w1_hi_bits, w1_lo_bits, b1, feat_exp, w_exp):
    """The tcgen05 stage's arithmetic with exact products and wide accumulation:
    leaky(2^-(fe+we) * (Ahi Bhi + Ahi Blo + Alo Bhi) + b1)."""
    ah, al = split_hi_lo(feat, feat_exp)
    bh = w1_hi_bits.view(np.float16).astype(np.float64).T     # [2688, 448]
    bl = w1_lo_bits.view(np.float16).astype(np.float64).T
    acc = ah.astype(np.float64) @ bh + ah.astype(np.float64) @ bl + al.astype(np.float64) @ bh
    return leaky((acc.astype(F) * F(2.0 ** -(feat_exp + w_exp)) + b1).astype(F))
