"""Lane-level numpy model of csrc/conv_stage.cu (test infrastructure).

It mirrors the kernel's work decomposition literally -- warp tasks, lane -> (quad, d) mapping, integer
pooled sums, the mma.sync.m16n8k16 fragment layouts (PTX ISA) chaining conv1 -> conv2 -> conv3 from
register to register, the packed weight-fragment layout of the shared-memory conv block, feature offsets
-- so that the index arithmetic of the CUDA code can be checked against the oracle on a box without a
GPU.  Arithmetic is done in float64 on un-split values (the hi/lo three-pass scheme is validated
separately); what this model pins is WHERE every number comes from and goes to.  Any change to the
mapping in conv_stage.cu / model.cpp must be mirrored here."""
import numpy as np

kOffC3S, kOffC3M, kOffC3L, kOffC2S, kOffC2M, kOffC2L = 0, 512, 640, 672, 2208, 2592
# conv block of one branch in 32-bit words (csrc/kernels.h)
kHdrOff, kB1Off, kB2Off, kB3Off = 0, 16, 32, 56
kF1HiOff, kF1LoOff, kF2HiOff, kF2LoOff, kF3HiOff, kF3LoOff, kW1SumOff, kBranchWords = 96, 224, 352, 1120, 1888, 3424, 4960, 4976
F = np.float32


def leaky(v):
    return np.maximum(0.2 * v, v)


def unpack_frags(block_words, off_hi, off_lo, n_frag, scale_exp):
    """[n_frag][32 lanes][2 regs][2 halves] -> float64 values (hi + lo) / 2^scale_exp."""
    hi = block_words[off_hi:off_hi + n_frag * 64].view(np.float16).astype(np.float64).reshape(n_frag, 32, 2, 2)
    lo = block_words[off_lo:off_lo + n_frag * 64].view(np.float16).astype(np.float64).reshape(n_frag, 32, 2, 2)
    return (hi + lo) / 2.0 ** scale_exp


def mma_m16n8k16(a_regs, b_regs, c_regs):
    """a_regs [32][4][2], b_regs [32][2][2], c_regs [32][4]: D = A(16x16) B(16x8) + C with the PTX fragment layouts."""
    A = np.zeros((16, 16))
    B = np.zeros((16, 8))
    C = np.zeros((16, 8))
    for lane in range(32):
        g, d = lane >> 2, lane & 3
        for e in range(2):
            A[g, 2 * d + e] = a_regs[lane][0][e]
            A[g + 8, 2 * d + e] = a_regs[lane][1][e]
            A[g, 2 * d + 8 + e] = a_regs[lane][2][e]
            A[g + 8, 2 * d + 8 + e] = a_regs[lane][3][e]
            B[2 * d + e, g] = b_regs[lane][0][e]
            B[2 * d + 8 + e, g] = b_regs[lane][1][e]
            C[g, 2 * d + e] = c_regs[lane][e]
            C[g + 8, 2 * d + e] = c_regs[lane][2 + e]
    D = A @ B + C
    out = np.zeros((32, 4))
    for lane in range(32):
        g, d = lane >> 2, lane & 3
        out[lane] = [D[g, 2 * d], D[g, 2 * d + 1], D[g + 8, 2 * d], D[g + 8, 2 * d + 1]]
    return out


def task_lane_info(task, lane):
    """(pool, branch, per-set (ctu, quad_y, quad_x), feature bases) of a lane, as the kernel computes them."""
    g, d = lane >> 2, lane & 3
    if task < 16:        # S: one CTU, 16 quads: set A = quads 0..7 (upper half), set B = quads 8..15
        P, br, G = 1, 0, 8
        sets = [(task, g >> 2, g & 3), (task, 2 + (g >> 2), g & 3)]
        c2_base, c3_base, QG = kOffC2S, kOffC3S, 4
    elif task < 20:      # M: four CTUs, 4 quads each: lanes 0..15 -> CTU 4t + 0 / +2, lanes 16..31 -> +1 / +3
        P, br, G = 2, 1, 4
        ca = 4 * (task - 16) + (g >> 2)
        q = g & 3
        sets = [(ca, q >> 1, q & 1), (ca + 2, q >> 1, q & 1)]
        c2_base, c3_base, QG = kOffC2M, kOffC3M, 2
    else:                # L: sixteen CTUs, one quad each
        P, br, G = 4, 2, 2
        sets = [(g, 0, 0), (g + 8, 0, 0)]
        c2_base, c3_base, QG = kOffC2L, kOffC3L, 1
    return P, br, G, QG, sets, c2_base, c3_base, g, d


def conv_features_group(tiles, conv_words, exps, cst3):
    """tiles [16,64,64] uint8; conv_words [3][4976] uint32 (the packed conv blocks); exps [3][4] = per branch
    (e1w, e_c1, e2w, e3w); cst3 = input_scale / (256 pool^2) per branch.  Returns float [16, 2688]."""
    feat = np.full((16, 2688), np.nan)
    # the producer warp's table: integer sums of the 16 blocks of 16x16 pixels of every tile, [ctu][4 qy + qx];
    # lane l adds chunk l%4 (16 bytes) of rows 8 jj + l/4, jj -> block row jj/2, then the 8 lanes of a chunk are reduced
    sums = np.zeros((16, 16), np.int64)
    for c in range(16):
        acc = np.zeros((32, 4), np.int64)
        for lane in range(32):
            for jj in range(8):
                row, ch = 8 * jj + (lane >> 2), lane & 3
                acc[lane, jj >> 1] += int(tiles[c][row, 16 * ch:16 * ch + 16].astype(np.int64).sum())
        for sh in (4, 8, 16):
            acc = acc + acc[np.arange(32) ^ sh]
        for lane in range(4):
            for q in range(4):
                sums[c, q * 4 + lane] = acc[lane, q]
    for task in range(21):
        info = [task_lane_info(task, lane) for lane in range(32)]
        P, br, G, QG = info[0][0], info[0][1], info[0][2], info[0][3]
        blk = conv_words[br]
        e1w, e_c1, e2w, e3w = exps[br]
        bias = blk.view(np.float32)
        b1, b2, b3 = bias[kB1Off:kB1Off + 16], bias[kB2Off:kB2Off + 24], bias[kB3Off:kB3Off + 32]
        f1 = unpack_frags(blk, kF1HiOff, kF1LoOff, 2, e1w)           # [nt]
        w1sum = bias[kW1SumOff:kW1SumOff + 16].astype(np.float64) / 2.0 ** e1w   # per channel: sum over the 16 taps
        f2 = unpack_frags(blk, kF2HiOff, kF2LoOff, 12, e2w)          # [p*3 + nt]
        f3 = unpack_frags(blk, kF3HiOff, kF3LoOff, 24, e3w)          # [j*4 + nt]
        pairs = np.zeros((2, 32, 12, 2))                              # conv3 A operand: [set][lane][s = 3r + nt][e]
        for st in range(2):
            # ---- window sum over the quad's 16x16 pooled block from the block-sum table: S = the quad's own block,
            # M = lane d takes block (2 qy + d/2, 2 qx + d%2), L = lane d takes block row d; M / L reduce over d
            wsum = np.zeros(32, np.int64)
            for lane in range(32):
                c, qy, qx = info[lane][4][st]
                d = lane & 3
                if P == 1:
                    wsum[lane] = sums[c, 4 * qy + qx]
                elif P == 2:
                    wsum[lane] = sums[c, (2 * qy + (d >> 1)) * 4 + 2 * qx + (d & 1)]
                else:
                    wsum[lane] = sums[c, 4 * d:4 * d + 4].sum()
            if P != 1:
                wsum = wsum + wsum[np.arange(32) ^ 1]
                wsum = wsum + wsum[np.arange(32) ^ 2]
            centre = 128 * P * P
            for T in range(2):
                D2 = np.zeros((3, 32, 4))
                for p in range(4):
                    # conv1 A fragment: rows g / g+8 = regions 2T / 2T+1 (rx = 0 / 1), k = tap = 4 ky + kx
                    a1 = np.zeros((32, 4, 2))
                    for lane in range(32):
                        c, qy, qx = info[lane][4][st]
                        d = lane & 3
                        pooled = tiles[c][16 * P * qy:16 * P * (qy + 1), 16 * P * qx:16 * P * (qx + 1)].astype(np.int64)
                        pooled = pooled.reshape(16, P, 16, P).sum(axis=(1, 3))
                        for reg in range(4):
                            rx, ky = reg & 1, (d >> 1) + 2 * (reg >> 1)
                            y = 8 * T + 4 * (p >> 1) + ky
                            for e in range(2):
                                x = 8 * rx + 4 * (p & 1) + 2 * (d & 1) + e
                                a1[lane, reg, e] = float(pooled[y, x] - centre)      # exact in fp16: |.| <= 2048
                    a2 = np.zeros((32, 4, 2))
                    for nt in range(2):
                        D1 = mma_m16n8k16(a1, f1[nt], np.zeros((32, 4)))
                        for lane in range(32):
                            d = lane & 3
                            for i in range(4):
                                ch = 8 * nt + 2 * d + (i & 1)
                                # conv2 A fragment regs: [0] row g lo cols, [1] row g+8 lo cols, [2] row g hi cols, [3] row g+8 hi cols
                                # conv(256 s - W) = 256 conv(s - centre) + (256 centre - W) * sum(taps)
                                pre = 256 * cst3[br] * D1[lane, i] + float(256 * centre - wsum[lane]) * cst3[br] * w1sum[ch]
                                a2[lane, 2 * nt + (i >> 1), i & 1] = leaky(pre + b1[ch])
                    for nt in range(3):
                        D2[nt] = mma_m16n8k16(a2, f2[p * 3 + nt], D2[nt])
                for nt in range(3):
                    for lane in range(32):
                        c, qy, qx = info[lane][4][st]
                        d = lane & 3
                        c2_base = info[lane][5]
                        for i in range(4):
                            r = 2 * T + (i >> 1)                     # region inside the quad = conv3 tap
                            ch = 8 * nt + 2 * d + (i & 1)
                            v = leaky(D2[nt][lane, i] + b2[ch])
                            ry, rx = 2 * qy + (r >> 1), 2 * qx + (r & 1)
                            feat[c, c2_base + (ry * G + rx) * 24 + ch] = v
                            pairs[st, lane, 3 * r + nt, i & 1] = v
        # ---- conv3: rows g = set A quad, g+8 = set B quad; k = 24 tap + channel = 16 j + col
        D3 = np.zeros((4, 32, 4))
        for j in range(6):
            a3 = np.zeros((32, 4, 2))
            for lane in range(32):
                a3[lane, 0] = pairs[0, lane, 2 * j]
                a3[lane, 1] = pairs[1, lane, 2 * j]
                a3[lane, 2] = pairs[0, lane, 2 * j + 1]
                a3[lane, 3] = pairs[1, lane, 2 * j + 1]
            for nt in range(4):
                D3[nt] = mma_m16n8k16(a3, f3[j * 4 + nt], D3[nt])
        for nt in range(4):
            for lane in range(32):
                d = lane & 3
                c3_base = info[lane][6]
                for i in range(4):
                    c, qy, qx = info[lane][4][i >> 1]
                    ch = 8 * nt + 2 * d + (i & 1)
                    feat[c, c3_base + (qy * QG + qx) * 32 + ch] = leaky(D3[nt][lane, i] + b3[ch])
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
    return np.maximum(F(0.2) * (acc.astype(F) * F(2.0 ** -(feat_exp + w_exp)) + b1).astype(F),
                      (acc.astype(F) * F(2.0 ** -(feat_exp + w_exp)) + b1).astype(F))


def pack_conv_block_reference(w, branch_base, input_bound):
    """Python statement of the conv block layout model.cpp must produce for one branch (used to check the C++
    packer): returns (words uint32[4976], (e1w, e_c1, e2w, e3w))."""
    def v(i):
        return w["Variable" if branch_base + i == 0 else "Variable_%d" % (branch_base + i)]
    w1, b1, w2, b2, w3, b3 = v(0).reshape(16, 16), v(1), v(2).reshape(64, 24), v(3), v(4).reshape(96, 32), v(5)

    def pick(bound):
        return int(min(14, max(-14, np.floor(np.log2(32768.0 / bound)))))
    B1 = np.abs(w1).astype(np.float64).sum(0) * input_bound + np.abs(b1)
    e1w, e2w, e3w = pick(np.abs(w1).max()), pick(np.abs(w2).max()), pick(np.abs(w3).max())
    e_c1 = pick(B1.max())
    words = np.zeros(kBranchWords, np.uint32)
    fl = words.view(np.float32)
    fl[kB1Off:kB1Off + 16], fl[kB2Off:kB2Off + 24], fl[kB3Off:kB3Off + 32] = b1, b2, b3

    def frags(W, ksteps, ntiles, exp, off_hi, off_lo):
        hi = np.zeros((ksteps * ntiles, 32, 2, 2), np.float16)
        lo = np.zeros_like(hi)
        for j in range(ksteps):
            for nt in range(ntiles):
                for lane in range(32):
                    g, d = lane >> 2, lane & 3
                    for reg in range(2):
                        for e in range(2):
                            s = np.float32(W[16 * j + 2 * d + 8 * reg + e, 8 * nt + g]) * np.float32(2.0 ** exp)
                            h = np.float16(s)
                            hi[j * ntiles + nt, lane, reg, e] = h
                            lo[j * ntiles + nt, lane, reg, e] = np.float16(s - np.float32(h))
        n = ksteps * ntiles * 64
        words[off_hi:off_hi + n] = hi.reshape(-1).view(np.uint32)
        words[off_lo:off_lo + n] = lo.reshape(-1).view(np.uint32)
    frags(w1, 1, 2, e1w, kF1HiOff, kF1LoOff)
    s1 = (w1.astype(np.float32) * np.float32(2.0 ** e1w)).astype(np.float32)
    h1 = s1.astype(np.float16)
    l1 = (s1 - h1.astype(np.float32)).astype(np.float16)
    fl[kW1SumOff:kW1SumOff + 16] = (h1.astype(np.float64) + l1.astype(np.float64)).sum(0).astype(np.float32)
    frags(w2, 4, 3, e2w, kF2HiOff, kF2LoOff)
    frags(w3, 6, 4, e3w, kF3HiOff, kF3LoOff)
    return words, (e1w, e_c1, e2w, e3w)
