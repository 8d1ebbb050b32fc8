"""Lane-level numpy model of csrc/conv_stage.cu (test infrastructure).

It mirrors the kernel's work decomposition literally -- warp tasks, lane -> region mapping, integer
pooled sums, quad shuffles, conv3 reduce-scatter, feature offsets, hi/lo fp16 split -- so that the index
arithmetic of the CUDA code can be checked against the oracle on a box without a GPU.  Any change to the
mapping in conv_stage.cu must be mirrored here."""
import numpy as np

kOffC3S, kOffC3M, kOffC3L, kOffC2S, kOffC2M, kOffC2L = 0, 512, 640, 672, 2208, 2592
kW1Off, kB1Off, kW2Off, kB2Off, kW3Off, kW3Stride, kB3Off, kBr = 0, 256, 272, 1808, 1832, 772, 4920, 4952
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
        for ci in range(24):
            o = kW3Off + d * kW3Stride + ci * 32
            part = (c2[ci] * wb[o:o + 32] + part).astype(F)
        return c2, part
    return rsum, rest


def conv_features_group(tiles, convw, cst3):
    """tiles: [8,64,64] uint8 (one tile group).  convw: [3,4952] packed conv weights (S, M, L).
    Returns float32 [8, 2688] features computed task by task exactly like the kernel."""
    feat = np.full((8, 2688), np.nan, F)
    for task in range(21):
        lanes = []
        for lane in range(32):
            if task < 16:
                c, half = task >> 1, task & 1
                q, d = lane >> 2, lane & 3
                qy, qx = half * 2 + (q >> 2), q & 3
                ry, rx = 2 * qy + (d >> 1), 2 * qx + (d & 1)
                P, br, oy, ox = 1, 0, 8 * ry, 8 * rx
                c2_off, c3_off = kOffC2S + (ry * 8 + rx) * 24, kOffC3S + (qy * 4 + qx) * 32
            elif task < 20:
                c = 2 * (task - 16) + (lane >> 4)
                l16 = lane & 15
                q, d = l16 >> 2, l16 & 3
                qy, qx = q >> 1, q & 1
                ry, rx = 2 * qy + (d >> 1), 2 * qx + (d & 1)
                P, br, oy, ox = 2, 1, 16 * ry, 16 * rx
                c2_off, c3_off = kOffC2M + (ry * 4 + rx) * 24, kOffC3M + (qy * 2 + qx) * 32
            else:
                c, d = lane >> 2, lane & 3
                ry, rx = d >> 1, d & 1
                P, br, oy, ox = 4, 2, 32 * ry, 32 * rx
                c2_off, c3_off = kOffC2L + (ry * 2 + rx) * 24, kOffC3L
            rsum, rest = lane_program(tiles[c], oy, ox, P, convw[br], cst3[br], d)
            lanes.append(dict(c=c, d=d, rsum=rsum, rest=rest, c2_off=c2_off, c3_off=c3_off, wb=convw[br]))
        # window sum: two xor-shuffles inside the quad
        r1 = [lanes[l]["rsum"] + lanes[l ^ 1]["rsum"] for l in range(32)]
        wsum = [r1[l] + r1[l ^ 2] for l in range(32)]
        parts = []
        for l in range(32):
            c2, part = lanes[l]["rest"](wsum[l])
            feat[lanes[l]["c"], lanes[l]["c2_off"]:lanes[l]["c2_off"] + 24] = c2
            parts.append(part)
        # reduce-scatter: lane d ends with channels [8d, 8d+8)
        r16 = []
        for l in range(32):
            up2 = (lanes[l]["d"] & 2) != 0
            send_from_peer = parts[l ^ 2][16:] if up2 else parts[l ^ 2][:16]   # what the peer sends = the half I keep
            keep = parts[l][16:] if up2 else parts[l][:16]
            r16.append((keep + send_from_peer).astype(F))
        for l in range(32):
            d = lanes[l]["d"]
            up1 = (d & 1) != 0
            peer = r16[l ^ 1][8:] if up1 else r16[l ^ 1][:8]
            keep = r16[l][8:] if up1 else r16[l][:8]
            r8 = (keep + peer).astype(F)
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
