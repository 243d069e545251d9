"""Run tools/umma_probe.cu on the GPU: correctness vs torch.matmul and MMA issue-rate probe."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "nano-vllm-ascend_b200", "lib", "libumma_probe.so"))
lib.umma_probe.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
lib.umma_rate.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p]
torch.manual_seed(0)
for (N, K, mode) in [(64, 128, 0), (128, 64, 1), (128, 128, 0), (64, 128, 2), (128, 64, 3), (128, 128, 3)]:
    A = torch.randn(128, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") if mode % 2 == 0 else torch.randn(K, N, device="cuda")).bfloat16()
    out = torch.full((128, N), float("nan"), device="cuda")
    rc = lib.umma_probe(A.data_ptr(), B.data_ptr(), out.data_ptr(), N, K, mode, None)
    torch.cuda.synchronize()
    want = A.float() @ (B.float().T if mode % 2 == 0 else B.float())
    print(f"probe N={N} K={K} mode={mode} rc={rc} max_err={(out - want).abs().max().item():.4g}")
cyc = torch.zeros(148, dtype=torch.int64, device="cuda")
for blocks in (1, 148):
    # bmn: bit 0 = B MN-major, bit 1 = A from tensor memory (the P operand of an FA4-shaped PV)
    for (N, ks, bmn) in [(64, 8, 0), (128, 8, 0), (256, 8, 0), (128, 4, 1), (64, 4, 1), (128, 8, 2), (128, 8, 3),
                         (128, 4, 3), (64, 4, 3)]:
        reps = 200
        lib.umma_rate(cyc.data_ptr(), blocks, N, ks, bmn, reps, None)
        torch.cuda.synchronize()
        c = cyc[:blocks].float().mean().item()
        per = c / (reps * ks)
        print(f"rate blocks={blocks} M=128 N={N} ksteps={ks} b_mn={bmn}: {per:.1f} cyc/MMA  (math floor {128*N/256:.0f}) -> {128*N*16/per:.0f} MAC/clk/SM")
