"""Lane-level emulation (numpy, CPU) of the data movement in paged_decode_fp8_kernel
(nano-vllm-ascend_b200/csrc/decode_attn_fp8.cu) for ONE consumer warp and ONE 16-token unit.

It models, with the PTX-documented register layouts,
  * the shared-memory image TMA produces for a {128 B, rows, chunks} box with the 128-byte swizzle,
  * ld.shared.v4 / ldmatrix.x4.trans / prmt as the kernel issues them,
  * mma.sync.m16n8k16 (A row-major, B col-major) assembled from the per-lane fragments,
and checks that the permuted contraction index of QK^T and the even/odd column split of PV
reproduce plain  S = Q K^T  and  O = P V. It guards the index arithmetic only (the part that
cannot be seen in a compiler listing); numerics are covered by the GPU parity tests.

    python tools/emulate_fp8_fragments.py        # prints "ok" per configuration
"""
from __future__ import annotations

import numpy as np


def swizzled_panel(rows_bytes: np.ndarray) -> np.ndarray:
    """rows_bytes [R, 128] (R % 8 == 0, panel base 1024-aligned) -> flat smem image."""
    R = rows_bytes.shape[0]
    img = np.zeros(R * 128, dtype=rows_bytes.dtype)
    for r in range(R):
        for c16 in range(8):
            phys = r * 128 + ((c16 ^ (r & 7)) << 4)
            img[phys:phys + 16] = rows_bytes[r, c16 * 16:(c16 + 1) * 16]
    return img


def mma_m16n8k16(a_regs, b_regs):
    """a_regs[lane] = (a0, a1, a2, a3) each a pair of values; b_regs[lane] = (b0, b1) pairs.
    Returns c[lane] = (c0, c1, c2, c3)."""
    A = np.zeros((16, 16))
    B = np.zeros((16, 8))
    for lane in range(32):
        g, t4 = lane >> 2, lane & 3
        a0, a1, a2, a3 = a_regs[lane]
        A[g, 2 * t4:2 * t4 + 2] = a0
        A[g + 8, 2 * t4:2 * t4 + 2] = a1
        A[g, 2 * t4 + 8:2 * t4 + 10] = a2
        A[g + 8, 2 * t4 + 8:2 * t4 + 10] = a3
        b0, b1 = b_regs[lane]
        B[2 * t4:2 * t4 + 2, g] = b0
        B[2 * t4 + 8:2 * t4 + 10, g] = b1
    C = A @ B
    out = []
    for lane in range(32):
        g, t4 = lane >> 2, lane & 3
        out.append((C[g, 2 * t4], C[g, 2 * t4 + 1], C[g + 8, 2 * t4], C[g + 8, 2 * t4 + 1]))
    return out


def ldmatrix_x4_trans(img: np.ndarray, lane_addr):
    """img: flat byte image; lane_addr[lane] = byte address of the 16-byte row this lane supplies
    (lanes 8m..8m+7 -> rows 0..7 of matrix m). Returns regs[lane][m] = 4 bytes (lo16 | hi16)."""
    mats = []
    for m in range(4):
        M = np.zeros((8, 8, 2), dtype=img.dtype)          # [row][b16 col][byte]
        for r in range(8):
            a = lane_addr[8 * m + r]
            M[r] = img[a:a + 16].reshape(8, 2)
        mats.append(M)
    regs = []
    for lane in range(32):
        g, t4 = lane >> 2, lane & 3
        # .trans: thread holds elements [row 2t4][col g] (low half) and [row 2t4+1][col g] (high half)
        regs.append([np.concatenate([mats[m][2 * t4, g], mats[m][2 * t4 + 1, g]]) for m in range(4)])
    return regs


def byte_perm(x, y, sel):
    src = np.concatenate([x, y])
    return np.array([src[(sel >> (4 * i)) & 0xF] for i in range(4)], dtype=x.dtype)


def run(D: int, head_local: int, G: int, valid: int = 16, seed: int = 0):
    rng = np.random.default_rng(seed)
    KSTEPS, NC = D // 16, D // 16
    head_byte = head_local * D
    hb16 = (head_byte & 127) >> 4
    # values stand in for de-quantised e4m3 numbers: small integers keep every product exact
    K = rng.integers(-3, 4, size=(16, D)).astype(np.float64)
    V = rng.integers(-3, 4, size=(16, D)).astype(np.float64)
    Q = rng.integers(-3, 4, size=(G, D)).astype(np.float64)
    # one 128-byte chunk row per token: this head's D bytes at offset head_byte & 127
    krows = np.zeros((16, 128))
    vrows = np.zeros((16, 128))
    krows[:, (head_byte & 127):(head_byte & 127) + D] = K
    vrows[:, (head_byte & 127):(head_byte & 127) + D] = V
    kimg, vimg = swizzled_panel(krows), swizzled_panel(vrows)

    # ---- QK^T
    S = np.zeros((G, 16))
    for n in range(2):
        acc = [np.zeros(4) for _ in range(32)]
        words = {}
        for lane in range(32):
            g, t4 = lane >> 2, lane & 3
            k_c16 = hb16 + t4 * (KSTEPS // 4)
            rowaddr = (n * 8 + g) * 128
            w = []
            for i in range(KSTEPS // 4):
                a = rowaddr + (((k_c16 + i) ^ g) << 4)
                w.extend(kimg[a:a + 16].reshape(4, 4))
            words[lane] = w                                   # words[lane][ks] = 4 bytes
        for ks in range(KSTEPS):
            a_regs, b_regs = [], []
            for lane in range(32):
                g, t4 = lane >> 2, lane & 3
                if g < G:
                    q4 = Q[g, t4 * (D // 4) + 4 * ks: t4 * (D // 4) + 4 * ks + 4]
                else:
                    q4 = np.zeros(4)
                a_regs.append((q4[0:2], np.zeros(2), q4[2:4], np.zeros(2)))
                w = words[lane][ks]
                b_regs.append((w[0:2], w[2:4]))
            c = mma_m16n8k16(a_regs, b_regs)
            for lane in range(32):
                acc[lane] += np.array(c[lane])
        for lane in range(32):
            g, t4 = lane >> 2, lane & 3
            if g < G:
                S[g, n * 8 + 2 * t4] = acc[lane][0]
                S[g, n * 8 + 2 * t4 + 1] = acc[lane][1]
    assert np.array_equal(S, Q @ K.T), "QK^T fragment map is wrong"

    # ---- PV with P := a known matrix (stands in for the softmax numerators)
    P = rng.integers(0, 3, size=(G, 16)).astype(np.float64)
    P[:, valid:] = 0
    # the kernel zeroes this head's bytes of V rows >= valid
    for idx in range((16 - valid) * NC):
        row = valid + idx // NC
        c = hb16 + idx % NC
        a = row * 128 + ((c ^ (row & 7)) << 4)
        vimg[a:a + 16] = 0
    o_acc = [[np.zeros(4) for _ in range(D // 8)] for _ in range(32)]
    for jp in range(D // 32):
        lane_addr = []
        for lane in range(32):
            v_tok = (((lane >> 3) & 1) << 3) + (lane & 7)
            v_dc = lane >> 4
            lane_addr.append(v_tok * 128 + (((hb16 + 2 * jp + v_dc) ^ (v_tok & 7)) << 4))
        regs = ldmatrix_x4_trans(vimg, lane_addr)
        frag = {t: [] for t in range(4)}
        a_regs = []
        for lane in range(32):
            g, t4 = lane >> 2, lane & 3
            r0, r1, r2, r3 = regs[lane]
            e0, o0 = byte_perm(r0, r1, 0x6420), byte_perm(r0, r1, 0x7531)
            e1, o1 = byte_perm(r2, r3, 0x6420), byte_perm(r2, r3, 0x7531)
            for t, x in enumerate((e0, o0, e1, o1)):
                frag[t].append((x[0:2], x[2:4]))
            if g < G:
                a_regs.append((P[g, 2 * t4:2 * t4 + 2], np.zeros(2), P[g, 2 * t4 + 8:2 * t4 + 10], np.zeros(2)))
            else:
                a_regs.append((np.zeros(2),) * 4)
        for t in range(4):
            c = mma_m16n8k16(a_regs, frag[t])
            for lane in range(32):
                o_acc[lane][4 * jp + t] += np.array(c[lane])
    O = np.zeros((G, D))
    for lane in range(32):
        g, t4 = lane >> 2, lane & 3
        if g >= G:
            continue
        for c in range(NC):
            O[g, 16 * c + 4 * t4: 16 * c + 4 * t4 + 4] = (o_acc[lane][2 * c][0], o_acc[lane][2 * c + 1][0],
                                                          o_acc[lane][2 * c][1], o_acc[lane][2 * c + 1][1])
    Vz = V.copy()
    Vz[valid:] = 0
    assert np.array_equal(O, P @ Vz), "PV fragment map is wrong"


if __name__ == "__main__":
    for D, heads in ((128, (0, 1, 5)), (64, (0, 1, 2, 3))):
        for h in heads:
            for G in (1, 2, 8):
                for valid in (16, 9, 1):
                    run(D, h, G, valid, seed=D + h + G + valid)
            print(f"ok D={D} head_local={h}")
