"""First-run diagnostics + kernel-only timing for the sparse-MoE kernels (csrc/moe.cu, csrc/moe_umma.cu).

  python tools/moe_diag.py check            # every stage against an fp32 torch evaluation on the GPU,
                                            # with an error MAP (which row / column blocks are wrong)
  python tools/moe_diag.py bench            # CUDA-event timing: decode shape (HBM GB/s of expert weights)
                                            # and prefill shape (TFLOP/s), streaming vs umma vs library
  ncu --set full -k regex:moe_ -c 6 python tools/moe_diag.py bench --once     # capture for profiles/

`check` is written for the FIRST hardware run of these kernels: instead of a bare assert it prints, per
stage, the max/mean error and the fraction of wrong elements per 16-row and per 64-column block, so that
an indexing mistake (a swapped fragment, a wrong tile offset, a missing row mask) is recognisable from a
single run.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nano-vllm-ascend_b200"))
from nanovllm import _C, ops  # noqa: E402

BF = torch.bfloat16


def make_case(T, k, E, hidden, inter, seed=0, empty_expert=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, hidden, generator=g).to(BF).cuda()
    w13 = (hidden ** -0.5 * torch.randn(E, 2 * inter, hidden, generator=g)).to(BF).cuda()
    w2 = (inter ** -0.5 * torch.randn(E, hidden, inter, generator=g)).to(BF).cuda()
    ids = torch.stack([torch.randperm(E, generator=g)[:k] for _ in range(T)]).to(torch.int32)
    if empty_expert and E > 2 and k == 1:
        ids[ids == 1] = 2                         # expert 1 receives nothing
    w = torch.rand(T, k, generator=g)
    w = (w / w.sum(-1, keepdim=True)).to(BF).cuda()
    return x, w13, w2, ids.cuda(), w


def error_map(name, got, want, row_block=16, col_block=64, tol=None):
    got, want = got.float(), want.float()
    d = (got - want).abs()
    scale = want.abs().max().item() + 1e-9
    tol = tol if tol is not None else 3e-2 * scale
    bad = d > tol
    print(f"[{name}] shape {tuple(got.shape)} max|ref| {scale:.4g} max err {d.max().item():.4g} "
          f"mean err {d.mean().item():.3g} wrong {bad.float().mean().item():.4%} (tol {tol:.3g}) "
          f"nan {int(torch.isnan(got).sum())}")
    if bad.any():
        R, C = got.shape
        rb = bad.view(-1).new_zeros(((R + row_block - 1) // row_block,), dtype=torch.float32)
        rows_bad = bad.float().mean(1)
        for i in range(rb.numel()):
            rb[i] = rows_bad[i * row_block:(i + 1) * row_block].mean()
        cols_bad = bad.float().mean(0)
        cb = torch.stack([cols_bad[j:j + col_block].mean() for j in range(0, C, col_block)])
        print(f"   wrong fraction per {row_block}-row block (first 32): {[round(v, 2) for v in rb[:32].tolist()]}")
        print(f"   wrong fraction per {col_block}-col block: {[round(v, 2) for v in cb.tolist()]}")
        idx = bad.nonzero()[:6].tolist()
        print("   first mismatches (row, col, got, want):",
              [(r, c, round(got[r, c].item(), 4), round(want[r, c].item(), 4)) for r, c in idx])
    return not bad.any()


def stage_reference(x, w13, w2, ids, w, plan):
    """fp32 torch evaluation of every intermediate buffer in PLAN order (rounding points of moe.cu)."""
    T, hidden = x.shape
    E, two_i, _ = w13.shape
    inter, k = two_i // 2, ids.shape[1]
    off = plan.offsets.cpu().tolist()
    sp = plan.sorted_pair.long()
    n = off[-1]
    tok = (sp[:n] // k)
    h = torch.zeros(T * k, inter, dtype=BF, device="cuda")
    y = torch.zeros(T * k, hidden, dtype=BF, device="cuda")
    for e in range(E):
        a, b = off[e], off[e + 1]
        if a == b:
            continue
        gu = (x[tok[a:b]].float() @ w13[e].float().T).to(BF)
        h[a:b] = torch.nn.functional.silu(gu[:, :inter]) * gu[:, inter:]
        y[a:b] = (h[a:b].float() @ w2[e].float().T).to(BF) * w.reshape(-1)[sp[a:b]].unsqueeze(1)
    out = torch.zeros(T, hidden, dtype=BF, device="cuda")
    pos = plan.pos_of_pair.view(T, k).long().cpu()
    idc = ids.cpu()
    for t in range(T):
        acc = None
        for e, j in sorted((int(idc[t, j]), j) for j in range(k)):
            if pos[t, j] < 0:
                continue
            v = y[pos[t, j]].float()
            acc = v if acc is None else (acc + v).to(BF).float()
        if acc is not None:
            out[t] = acc.to(BF)
    return h, y, out, n


def run_stages(x, w13, w2, ids, w, plan, path, h_override=None):
    """The product kernels stage by stage (path: 'stream' or 'umma'); returns h, y, out.
    h_override: feed this H to the second GEMM instead of the first GEMM's output (isolates stage 2)."""
    lib = _C.load()
    T, hidden = x.shape
    n_local, two_i, _ = w13.shape
    inter, k = two_i // 2, plan.top_k
    pairs = T * k
    dt, st = ops._dt(x), ops._stream()
    h = torch.zeros(pairs, inter, dtype=BF, device="cuda")
    y = torch.zeros(pairs, hidden, dtype=BF, device="cuda")
    h2 = h if h_override is None else h_override.contiguous()
    if path == "stream":
        _C.check(lib.b200_moe_gemm_gate_up(h.data_ptr(), x.data_ptr(), w13.data_ptr(), plan.offsets.data_ptr(),
                                           plan.sorted_pair.data_ptr(), hidden, inter, n_local, k, x.stride(0),
                                           dt, st), "gate_up")
        _C.check(lib.b200_moe_gemm_down(y.data_ptr(), h2.data_ptr(), w2.data_ptr(), plan.offsets.data_ptr(),
                                        plan.sorted_pair.data_ptr(), w.data_ptr(), inter, hidden, n_local, k,
                                        dt, st), "down")
    else:
        xs = torch.zeros(pairs, hidden, dtype=BF, device="cuda")
        _C.check(lib.b200_moe_gather_rows(xs.data_ptr(), x.data_ptr(), plan.offsets.data_ptr(),
                                          plan.sorted_pair.data_ptr(), pairs, n_local, k, hidden, x.stride(0),
                                          dt, st), "gather")
        _C.check(lib.b200_moe_gemm_gate_up_umma(h.data_ptr(), xs.data_ptr(), w13.data_ptr(),
                                                plan.offsets.data_ptr(), pairs, hidden, inter, n_local, dt, st),
                 "gate_up_umma")
        _C.check(lib.b200_moe_gemm_down_umma(y.data_ptr(), h2.data_ptr(), w2.data_ptr(), plan.offsets.data_ptr(),
                                             plan.sorted_pair.data_ptr(), w.data_ptr(), pairs, inter, hidden,
                                             n_local, dt, st), "down_umma")
    out = torch.empty(T, hidden, dtype=BF, device="cuda")
    _C.check(lib.b200_moe_combine(out.data_ptr(), y.data_ptr(), ids.data_ptr(), plan.pos_of_pair.data_ptr(), T, k,
                                  hidden, hidden, dt, st), "combine")
    torch.cuda.synchronize()
    return h, y, out


def check():
    ok = True
    cases = [("decode-like", dict(T=256, k=8, E=128, hidden=512, inter=256)),
             ("30b-shapes", dict(T=64, k=8, E=128, hidden=2048, inter=768)),
             ("many-rows", dict(T=700, k=2, E=4, hidden=256, inter=128)),
             ("ragged-top1", dict(T=97, k=1, E=3, hidden=128, inter=64, empty_expert=True))]
    for name, kw in cases:
        x, w13, w2, ids, w = make_case(**kw)
        E = kw["E"]
        for begin, n in ((0, E), (E // 2, E - E // 2)):
            plan = ops.moe_align(ids, begin, n)
            torch.cuda.synchronize()
            wa, wb = w13[begin:begin + n].contiguous(), w2[begin:begin + n].contiguous()
            rh, ry, rout, npairs = stage_reference(x, wa, wb, ids, w, plan)
            for path in ("stream", "umma"):
                tag = f"{name} experts[{begin},{begin + n}) {path}"
                try:
                    h, y, out = run_stages(x, wa, wb, ids, w, plan, path)
                except Exception as e:            # keep going: the other path may be fine
                    print(f"[{tag}] FAILED to run: {type(e).__name__}: {e}")
                    ok = False
                    continue
                h_ok = error_map(tag + " H", h[:npairs], rh[:npairs])
                if not h_ok:      # a wrong first stage must not mask the second: redo it on the reference H
                    h, y, out = run_stages(x, wa, wb, ids, w, plan, path, h_override=rh)
                ok &= h_ok
                ok &= error_map(tag + " Y", y[:npairs], ry[:npairs])
                ok &= error_map(tag + " out", out, rout)
    print("ALL OK" if ok else "MISMATCHES (see maps above)")
    return ok


def timeit(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3      # us


def bench(once=False):
    iters, warm = (1, 1) if once else (20, 5)
    hidden, inter, E, k = 2048, 768, 128, 8            # Qwen3-30B-A3B
    # 1.2 GB of expert weights per call >> 126 MB of L2: every iteration streams them from HBM
    for T, label in ((256, "decode B=256"), (4096, "prefill 4096 tokens")):
        x, w13, w2, ids, w = make_case(T, k, E, hidden, inter)
        plan = ops.moe_align(ids, 0, E)
        hit = int((plan.offsets[1:] > plan.offsets[:-1]).sum())
        wbytes = hit * 3 * hidden * inter * 2
        flops = 2 * T * k * 3 * hidden * inter
        for path in (("stream",) if T <= 512 else ("stream", "umma", "library")):
            prev = (ops._MOE_BIG_GEMM, ops._MOE_BIG_GEMM_ROWS)
            ops._MOE_BIG_GEMM, ops._MOE_BIG_GEMM_ROWS = path, 1
            try:
                us = timeit(lambda: ops.moe_experts(x, w13, w2, ids, w, plan), iters, warm)
                print(f"{label:22s} {path:8s} {us:9.1f} us  expert-weight stream {wbytes / us / 1e3:8.1f} GB/s  "
                      f"{flops / us / 1e6:8.1f} TFLOP/s  (experts hit {hit}/{E})")
            except Exception as e:
                print(f"{label:22s} {path:8s} FAILED: {type(e).__name__}: {e}")
            finally:
                ops._MOE_BIG_GEMM, ops._MOE_BIG_GEMM_ROWS = prev
        logits = torch.randn(T, E, device="cuda").to(BF)
        us_r = timeit(lambda: ops.moe_route(logits, k), iters, warm)
        us_a = timeit(lambda: ops.moe_align(ids, 0, E), iters, warm)
        print(f"{label:22s} route {us_r:7.1f} us   align (2 kernels + plan buffers) {us_a:7.1f} us")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["check", "bench"])
    ap.add_argument("--once", action="store_true", help="single iteration (for ncu captures)")
    a = ap.parse_args()
    if a.what == "check":
        sys.exit(0 if check() else 1)
    bench(a.once)
