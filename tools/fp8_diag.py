"""GPU diagnostic for the fp8 KV kernels (not a test): isolates the QK^T and PV fragment maps.
  python tools/fp8_diag.py        # on a B200; prints max-abs errors per sub-case
Reference values are computed with torch on the GPU itself (fp32), no oracle import."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nano-vllm-ascend_b200")]
import torch  # noqa: E402
from nanovllm import ops  # noqa: E402


def deq(b):
    return b.view(torch.float8_e4m3fn).float()


def ref(q, kc, vc, bt, ctx, G, D):
    B, H, _ = q.shape
    out = torch.zeros(B, H, D, device=q.device)
    for i in range(B):
        n = int(ctx[i])
        if n == 0:
            continue
        ids = bt[i][bt[i] >= 0].long()
        k = deq(kc[ids]).view(-1, G, D)[:n]
        v = deq(vc[ids]).view(-1, G, D)[:n]
        k = k.repeat_interleave(H // G, 1)
        v = v.repeat_interleave(H // G, 1)
        s = torch.einsum("hd,thd->ht", q[i].float(), k) * D ** -0.5
        out[i] = torch.einsum("ht,thd->hd", s.softmax(-1), v)
    return out.view(B, H * D)


def case(name, H, G, D, bs, lens, k_mode="rand", v_mode="rand", q_mode="rand"):
    torch.manual_seed(0)
    B = len(lens)
    nblk = [(n + bs - 1) // bs for n in lens]
    nb = sum(nblk) + 2
    perm = torch.randperm(nb, device="cuda")
    bt = torch.full((B, max(nblk)), -1, dtype=torch.int32, device="cuda")
    p = 0
    for i, n in enumerate(nblk):
        bt[i, :n] = perm[p:p + n].int()
        p += n
    shape = (nb, bs, G * D)
    kf = torch.randn(shape, device="cuda") * 0.5 if k_mode == "rand" else torch.zeros(shape, device="cuda")
    if v_mode == "rand":
        vf = torch.randn(shape, device="cuda")
    elif v_mode == "index":      # V[t][h][d] = d/16: exposes column permutations exactly
        vf = (torch.arange(G * D, device="cuda") % D / 16.0).expand(shape).contiguous()
    else:
        vf = torch.ones(shape, device="cuda")
    kc = kf.clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8)
    vc = vf.clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8)
    q = torch.randn(B, H, D, device="cuda").bfloat16() if q_mode == "rand" else torch.zeros(
        B, H, D, device="cuda", dtype=torch.bfloat16)
    ctx = torch.tensor(lens, dtype=torch.int32, device="cuda")
    t0 = time.time()
    o = ops.paged_decode_attn_fp8(q, kc, vc, bt, ctx, D ** -0.5, G)
    torch.cuda.synchronize()
    want = ref(q, kc, vc, bt, ctx, G, D)
    err = (o.float() - want).abs()
    print(f"{name:34s} H={H} G={G} D={D} bs={bs}: max err {err.max().item():.4f} "
          f"(rows {err.view(B, -1).max(1).values.tolist()}) {time.time() - t0:.2f}s", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    case("uniform P, V=index (PV col map)", 16, 8, 128, 16, [16], k_mode="zero", v_mode="index", q_mode="zero")
    case("V=1 (softmax sums to 1)", 16, 8, 128, 16, [16, 40])
    case("full, one block", 16, 8, 128, 16, [16])
    case("full, ragged", 16, 8, 128, 16, [1, 17, 100, 1024, 0, 300])
    case("D=64", 32, 8, 64, 16, [16, 100, 333])
    case("Hkv=1 (8 warps/head)", 8, 1, 128, 16, [16, 200, 1000])
    case("block 256", 16, 8, 128, 256, [700, 30])
    # bandwidth: config-2 shape
    B, H, G, D, bs, n = 256, 16, 8, 128, 16, 1024
    nb = B * n // bs
    kc = torch.randint(0, 120, (nb + 1, bs, G * D), dtype=torch.uint8, device="cuda")
    vc = torch.randint(0, 120, (nb + 1, bs, G * D), dtype=torch.uint8, device="cuda")
    q = torch.randn(B, H, D, device="cuda").bfloat16()
    bt = torch.randperm(nb, device="cuda").int().view(B, n // bs)
    ctx = torch.full((B,), n, dtype=torch.int32, device="cuda")
    o = torch.empty(B, H * D, device="cuda", dtype=torch.bfloat16)
    ws = ops.decode_workspace(B, H, G, D, q.device)
    for _ in range(3):
        ops.paged_decode_attn_fp8(q, kc, vc, bt, ctx, D ** -0.5, G, out=o, workspace=ws)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.paged_decode_attn_fp8(q, kc, vc, bt, ctx, D ** -0.5, G, out=o, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"fp8 decode B=256 ctx=1024: {us:.1f} us/launch, {B * n * 2 * G * D / us / 1e3:.0f} GB/s "
          f"(K+V cache {2 * nb * bs * G * D / 2**20:.0f} MiB)", flush=True)
