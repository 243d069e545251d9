"""Thread-level emulation (numpy, CPU) of the index arithmetic in nano-vllm-ascend_b200/csrc/moe.cu.

It transcribes, formula by formula,
  * moe_count_kernel / moe_place_kernel (per-warp chunks, ballot ranks) and checks the result
    against oracle/moe_ref.py::dispatch_plan,
  * moe_gemm_pass<gate_up / down, MI = 1 / 2>: the cp.async placement of every thread into the
    128-byte-swizzled stage, the ldmatrix.x4 row addresses of every lane, mma.sync.m16n8k16
    assembled from the PTX-documented fragment layouts, and the epilogue's (row, column)
    mapping -- and checks  H = silu(X W_gate^T) * (X W_up^T)  /  Y = (H W2^T) * scale  on
    integer-valued data (exact in fp64), including ragged row counts and the 64-row + 32-row
    pass split,
  * moe_combine_kernel's ascending-expert insertion sort.
It guards what a compiler listing cannot show (who loads what, where); numerics and the real
instruction behaviour are covered by the GPU parity tests (tests/test_z_staged_gpu.py).

    python tools/emulate_moe_kernels.py        # prints "ok ..." per configuration
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import moe_ref  # noqa: E402


# ------------------------------------------------------------------ dispatch plan
def emulate_align(ids_flat: np.ndarray, expert_begin: int, num_local: int):
    P = len(ids_flat)
    counts = np.zeros(num_local, np.int64)
    pos_of_pair = np.full(P, 12345, np.int64)              # garbage: the count kernel must reset it
    grid = num_local
    for e in range(grid):                                  # moe_count_kernel
        per = (P + grid - 1) // grid
        hi = min(P, (e + 1) * per)
        for tid in range(256):
            for i in range(e * per + tid, hi, 256):
                pos_of_pair[i] = -1
        counts[e] = sum(int(ids_flat[i] == expert_begin + e) for i in range(P))
    offsets = np.zeros(num_local + 1, np.int64)
    sorted_pair = np.full(P, -7, np.int64)
    for e in range(grid):                                  # moe_place_kernel
        target = expert_begin + e
        s_base = int(counts[:e].sum())
        chunk = (((P + 7) // 8) + 31) // 32 * 32
        warp_cnt = []
        for warp in range(8):
            lo = min(P, warp * chunk)
            hi = min(P, lo + chunk)
            warp_cnt.append(sum(int(ids_flat[i] == target) for i in range(lo, hi)))
        for warp in range(8):
            lo = min(P, warp * chunk)
            hi = min(P, lo + chunk)
            before = s_base + sum(warp_cnt[:warp])
            for i0 in range(lo, hi, 32):
                ballot = [(i0 + lane < hi) and ids_flat[i0 + lane] == target for lane in range(32)]
                for lane in range(32):
                    if ballot[lane]:
                        pos = before + sum(ballot[:lane])
                        sorted_pair[pos] = i0 + lane
                        pos_of_pair[i0 + lane] = pos
                before += sum(ballot)
        offsets[e] = s_base
        if e == num_local - 1:
            offsets[num_local] = s_base + sum(warp_cnt)
    return offsets, sorted_pair, pos_of_pair


def check_align():
    rng = np.random.default_rng(0)
    for T, k, E, begin, n in ((37, 2, 8, 0, 8), (256, 8, 128, 0, 128), (256, 8, 128, 32, 16),
                              (1, 8, 128, 112, 16), (50, 4, 32, 8, 8), (300, 8, 128, 0, 16)):
        ids = np.stack([rng.permutation(E)[:k] for _ in range(T)]).astype(np.int32)
        off, sp, pos = emulate_align(ids.reshape(-1), begin, n)
        w_off, w_sp, w_pos = moe_ref.dispatch_plan(torch.from_numpy(ids), begin, n)
        nl = int(w_off[-1])
        assert np.array_equal(off, w_off.numpy()), (T, k, E)
        assert np.array_equal(sp[:nl], w_sp.numpy())
        assert np.array_equal(pos, w_pos.numpy())
        print(f"ok align T={T} k={k} E={E} local=[{begin},{begin + n}) pairs={nl}")


# ------------------------------------------------------------------ expert GEMM pass
def tile_off(row, c16):
    return row * 128 + ((c16 ^ (row & 7)) << 4)


def ldsm_x4(img, lane_addr):
    """Non-transposed ldmatrix.x4 over a flat image of 2-byte elements stored as one float per
    element (addresses are byte addresses; 16 bytes = 8 elements). Returns regs[lane][m] = the two
    elements thread `lane` receives from matrix m: row g, columns 2*t4 and 2*t4+1."""
    regs = []
    for lane in range(32):
        g, t4 = lane >> 2, lane & 3
        r = []
        for m in range(4):
            a = lane_addr[8 * m + g] // 2
            r.append(img[a + 2 * t4:a + 2 * t4 + 2].copy())
        regs.append(r)
    return regs


def mma(acc, a_regs, b0, b1):
    """acc[lane] (4,) += m16n8k16 of the fragments. a_regs[lane] = [a0..a3], b0/b1[lane] pairs."""
    A = np.zeros((16, 16))
    B = np.zeros((16, 8))
    for lane in range(32):
        g, t4 = lane >> 2, lane & 3
        a0, a1, a2, a3 = a_regs[lane]
        A[g, 2 * t4:2 * t4 + 2] = a0
        A[g + 8, 2 * t4:2 * t4 + 2] = a1
        A[g, 2 * t4 + 8:2 * t4 + 10] = a2
        A[g + 8, 2 * t4 + 8:2 * t4 + 10] = a3
        B[2 * t4:2 * t4 + 2, g] = b0[lane]
        B[2 * t4 + 8:2 * t4 + 10, g] = b1[lane]
    C = A @ B
    for lane in range(32):
        g, t4 = lane >> 2, lane & 3
        acc[lane] += np.array([C[g, 2 * t4], C[g, 2 * t4 + 1], C[g + 8, 2 * t4], C[g + 8, 2 * t4 + 1]])


def emulate_pass(gate_up, MI, a, a_stride_rows, w, out, sorted_pair, pair_scale, row_begin, cnt, m0, n0,
                 K, n_out, top_k):
    """a: [rows, K] matrix the pass gathers from; w: this expert's [2*n_out or n_out, K]."""
    AROWS, BROWS, BK = 64, 128, 64
    KB = K // BK
    acc = np.zeros((8, 2 * MI, 2, 32, 4))                  # [warp][mf][nf][lane][4]
    for kb in range(KB):
        sa = np.full(AROWS * 64, np.nan)                   # stage images, one float per element
        sb = np.full(BROWS * 64, np.nan)
        for tid in range(256):                             # load_stage
            c16, rowbase = tid & 7, tid >> 3
            for i in range(MI):
                r = m0 + rowbase + 32 * i
                ok = r < cnt
                pos = row_begin + (r if ok else 0)
                src_row = sorted_pair[pos] // top_k if gate_up else pos
                dst = tile_off(rowbase + 32 * i, c16) // 2
                sa[dst:dst + 8] = a[src_row, kb * BK + c16 * 8:kb * BK + c16 * 8 + 8] if ok else 0.0
            for i in range(4):
                r = rowbase + 32 * i
                wrow = (n0 + r if r < 64 else n_out + n0 + (r - 64)) if gate_up else n0 + r
                dst = tile_off(r, c16) // 2
                sb[dst:dst + 8] = w[wrow, kb * BK + c16 * 8:kb * BK + c16 * 8 + 8]
        for warp in range(8):
            for ks in range(BK // 16):
                baddr, aaddr = [], [[] for _ in range(2 * MI)]
                for lane in range(32):
                    a_row = (((lane >> 3) & 1) << 3) + (lane & 7)
                    a_half = lane >> 4
                    if gate_up:
                        b_row = (64 if (lane >> 4) else 0) + warp * 8 + (lane & 7)
                    else:
                        b_row = warp * 16 + ((lane >> 4) << 3) + (lane & 7)
                    b_half = (lane >> 3) & 1
                    baddr.append(tile_off(b_row, ks * 2 + b_half))
                    for mf in range(2 * MI):
                        aaddr[mf].append(tile_off(mf * 16 + a_row, ks * 2 + a_half))
                b = ldsm_x4(sb, baddr)
                for mf in range(2 * MI):
                    ar = ldsm_x4(sa, aaddr[mf])
                    mma(acc[warp][mf][0], ar, [b[l][0] for l in range(32)], [b[l][1] for l in range(32)])
                    mma(acc[warp][mf][1], ar, [b[l][2] for l in range(32)], [b[l][3] for l in range(32)])
    assert not np.isnan(acc).any(), "a fragment read shared memory nobody wrote"
    for warp in range(8):                                  # epilogue
        for lane in range(32):
            g, t4 = lane >> 2, lane & 3
            for mf in range(2 * MI):
                for r in range(2):
                    row = m0 + mf * 16 + g + 8 * r
                    if row >= cnt:
                        continue
                    pos = row_begin + row
                    c = acc[warp][mf]
                    if gate_up:
                        for j in range(2):
                            gv, uv = c[0][lane][2 * r + j], c[1][lane][2 * r + j]
                            col = n0 + warp * 8 + 2 * t4 + j
                            assert np.isnan(out[pos, col]), "written twice"
                            out[pos, col] = gv / (1 + np.exp(-gv)) * uv
                    else:
                        sc = pair_scale[sorted_pair[pos]]
                        for nf in range(2):
                            for j in range(2):
                                col = n0 + warp * 16 + nf * 8 + 2 * t4 + j
                                assert np.isnan(out[pos, col]), "written twice"
                                out[pos, col] = c[nf][lane][2 * r + j] * sc


def emulate_gemm(gate_up, a, w_all, offsets, sorted_pair, pair_scale, K, n_out, top_k, pairs):
    out = np.full((pairs, n_out), np.nan)
    num_local = len(offsets) - 1
    for e in range(num_local):
        for bx in range(n_out // (64 if gate_up else 128)):
            row_begin, cnt = int(offsets[e]), int(offsets[e + 1] - offsets[e])
            if cnt <= 0:
                continue
            n0 = bx * (64 if gate_up else 128)
            m0 = 0
            while m0 < cnt:
                if cnt - m0 > 32:
                    emulate_pass(gate_up, 2, a, None, w_all[e], out, sorted_pair, pair_scale, row_begin,
                                 cnt, m0, n0, K, n_out, top_k)
                    m0 += 64
                else:
                    emulate_pass(gate_up, 1, a, None, w_all[e], out, sorted_pair, pair_scale, row_begin,
                                 cnt, m0, n0, K, n_out, top_k)
                    m0 += 32
    return out


def check_gemm():
    rng = np.random.default_rng(1)
    # (tokens, top_k, experts, hidden, inter): row counts per expert span 0, <32, 33..64, >64
    for T, k, E, hidden, inter in ((20, 2, 4, 128, 64), (90, 2, 2, 128, 64), (130, 1, 3, 128, 128)):
        ids = np.stack([rng.permutation(E)[:k] for _ in range(T)]).astype(np.int32)
        if E == 3:
            ids[:] = np.where(rng.random((T, 1)) < 0.75, 0, 2)       # expert 1 empty, expert 0 > 64 rows
        off, sp, pos = (t.numpy() for t in moe_ref.dispatch_plan(torch.from_numpy(ids), 0, E))
        P = T * k
        spad = np.concatenate([sp, np.zeros(P - len(sp), np.int32)])
        x = rng.integers(-2, 3, (T, hidden)).astype(np.float64)
        w13 = rng.integers(-2, 3, (E, 2 * inter, hidden)).astype(np.float64) / 8
        w2 = rng.integers(-2, 3, (E, hidden, inter)).astype(np.float64) / 8
        scale = rng.integers(1, 5, P).astype(np.float64) / 4
        h = emulate_gemm(True, x, w13, off, spad, None, hidden, inter, k, P)
        y = emulate_gemm(False, np.nan_to_num(h), w2, off, spad, scale, inter, hidden, k, P)
        for e in range(E):
            for p_ in range(int(off[e]), int(off[e + 1])):
                t = spad[p_] // k
                gu = w13[e] @ x[t]
                want_h = gu[:inter] / (1 + np.exp(-gu[:inter])) * gu[inter:]
                assert np.allclose(h[p_], want_h, rtol=1e-12, atol=1e-12), (e, p_)
                assert np.allclose(y[p_], (w2[e] @ want_h) * scale[spad[p_]], rtol=1e-9, atol=1e-9)
        assert np.isnan(h[int(off[-1]):]).all() and np.isnan(y[int(off[-1]):]).all()   # nothing beyond the plan
        counts = np.diff(off).tolist()
        print(f"ok gemm T={T} k={k} E={E} hidden={hidden} inter={inter} rows/expert={counts}")


def check_combine_order():
    rng = np.random.default_rng(2)
    for _ in range(200):
        k = int(rng.integers(1, 9))
        ids = rng.permutation(128)[:k]
        pos_in = np.where(rng.random(k) < 0.7, rng.integers(0, 1000, k), -1)
        s_ids, s_pos = [], []
        for j in range(k):                                  # the kernel's insertion sort
            if pos_in[j] < 0:
                continue
            at = len(s_ids)
            s_ids.append(None)
            s_pos.append(None)
            while at > 0 and s_ids[at - 1] > ids[j]:
                s_ids[at], s_pos[at] = s_ids[at - 1], s_pos[at - 1]
                at -= 1
            s_ids[at], s_pos[at] = ids[j], pos_in[j]
        want = sorted((int(i), int(p)) for i, p in zip(ids, pos_in) if p >= 0)
        assert [(int(i), int(p)) for i, p in zip(s_ids, s_pos)] == want
    print("ok combine order")


def check_route():
    """moe_route_kernel, lane by lane: expert e lives in lane e % 32, slot e // 32; k rounds of
    (per-lane best over slots, ties to the lower slot) -> (warp arg-max, ties to the lower index) ->
    the owning lane retires the winner. Against a stable descending sort of the same probabilities."""
    rng = np.random.default_rng(4)
    for E, k in ((8, 2), (5, 2), (33, 1), (48, 6), (128, 8), (256, 8)):
        VPL = (E + 31) // 32
        VPL = 1 if VPL <= 1 else 2 if VPL <= 2 else 4 if VPL <= 4 else 8
        for trial in range(40):
            # coarse values -> many exact ties
            logits = rng.integers(-3, 4, E).astype(np.float32) / 2
            pe = np.exp(logits - logits.max())
            probs = (pe / pe.sum()).astype(np.float32)
            p = np.full((32, VPL), -1.0, np.float32)
            for lane in range(32):
                for i in range(VPL):
                    e = lane + 32 * i
                    if e < E:
                        p[lane, i] = probs[e]
            sel = []
            for j in range(k):
                bv = np.full(32, -1.0, np.float32)
                bi = np.full(32, 0x7FFFFFFF, np.int64)
                for lane in range(32):
                    for i in range(VPL):
                        if p[lane, i] > bv[lane]:
                            bv[lane], bi[lane] = p[lane, i], lane + 32 * i
                o = 16
                while o > 0:                                   # butterfly arg-max
                    nv, ni = bv.copy(), bi.copy()
                    for lane in range(32):
                        ov, oi = bv[lane ^ o], bi[lane ^ o]
                        if ov > bv[lane] or (ov == bv[lane] and oi < bi[lane]):
                            nv[lane], ni[lane] = ov, oi
                    bv, bi = nv, ni
                    o >>= 1
                assert (bi == bi[0]).all()
                w = int(bi[0])
                p[w & 31, w >> 5] = -1.0
                sel.append(w)
            want = sorted(range(E), key=lambda e: (-probs[e], e))[:k]
            assert sel == want, (E, k, sel, want)
    print("ok route selection (ties -> lower index)")


def check_umma_tile_walk():
    """moe_umma_gemm_kernel: blockIdx.y -> (expert, 128-row tile) by walking the offsets; the grid's
    upper bound (pairs/128 + experts) must cover every tile, every row must be stored exactly once,
    and rows a tile computes past its expert's end must be discarded."""
    rng = np.random.default_rng(3)
    for trial in range(50):
        E = int(rng.integers(1, 20))
        counts = rng.integers(0, 400, E) * (rng.random(E) < 0.8)
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        pairs = int(offsets[-1] + rng.integers(0, 300))          # capacity >= local pairs
        grid_y = (pairs + 127) // 128 + E
        stored = np.zeros(pairs + 256, np.int64)
        for tile in range(grid_y):
            e, t = -1, 0
            for i in range(E):
                b, c = offsets[i], offsets[i + 1] - offsets[i]
                nt = (c + 127) >> 7
                if tile < t + nt:
                    e, m0, row_begin, cnt = i, (tile - t) << 7, b, c
                    break
                t += nt
            if e < 0:
                continue
            for row in range(128):
                if m0 + row < cnt:
                    stored[row_begin + m0 + row] += 1
        assert (stored[:offsets[-1]] == 1).all() and (stored[offsets[-1]:] == 0).all(), trial
        assert sum((c + 127) // 128 for c in counts) <= grid_y
    print("ok umma tile walk")


if __name__ == "__main__":
    check_route()
    check_umma_tile_walk()
    check_align()
    check_gemm()
    check_combine_order()
