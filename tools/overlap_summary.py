"""Summarise a chrome trace written by tools/overlap_timeline.py (one replayed decode step): per-kernel
launch counts and device times, the fraction of the step during which a paged-decode attention kernel is
resident (the HBM-bound part), and the kernels of one layer in start order with their streams.
Usage: python tools/overlap_summary.py gpurun_out/trace_<variant>.json [--window 2500 2720]"""
import argparse
import collections
import json


def short(name: str) -> str:
    for key, label in (("paged_decode_kernel", "paged_decode_kernel"), ("decode_combine", "decode_combine_kernel"),
                       ("qknorm_rope", "qknorm_rope_cache_kernel"), ("rmsnorm", "rmsnorm_kernel"),
                       ("silu_and_mul", "silu_and_mul_kernel"), ("sample_kernel", "sample_kernel"),
                       ("moe_umma", "moe_umma_gemm_kernel (dense form)")):
        if key in name:
            return label
    if "nvjet" in name:
        return "cuBLAS " + name.split("nvjet_tst_")[1].split("_v_")[0].split("_h_")[0]
    return name[:40]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("trace")
    ap.add_argument("--window", type=float, nargs=2, default=None, help="print every kernel starting in [a, b) us")
    args = ap.parse_args()
    with open(args.trace) as f:
        ev = [e for e in json.load(f)["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    span = max(e["ts"] + e["dur"] for e in ev) - t0
    print(f"# {args.trace}: {len(ev)} kernels, step span {span:.1f} us, streams "
          f"{sorted({e['args'].get('stream') for e in ev})}")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in ev:
        a = agg[short(e["name"])]
        a[0] += 1
        a[1] += e["dur"]
    print("| kernel | launches | total us | avg us |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {t:.1f} | {t / n:.2f} |")
    att = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev if "paged_decode_kernel" in e["name"])
    busy, (cs, ce) = 0.0, att[0]
    for s, e in att[1:]:
        if s > ce:
            busy += ce - cs
            cs, ce = s, e
        else:
            ce = max(ce, e)
    busy += ce - cs
    print(f"attention resident {busy:.1f} us = {busy / span:.3f} of the step "
          f"(sum of its launches {sum(e - s for s, e in att):.1f} us)")
    print("(durations of kernels launched with the PDL attribute include their wait for the predecessor)")
    if args.window:
        for e in ev:
            s = e["ts"] - t0
            if args.window[0] <= s < args.window[1]:
                a = e["args"]
                print(f"  stream {a.get('stream'):>4} {s:8.1f} +{e['dur']:6.1f} -> {s + e['dur']:8.1f}  "
                      f"{short(e['name']):34s} grid={a.get('grid')} smem={a.get('shared memory')}")


if __name__ == "__main__":
    main()
