"""Developer diagnostics for the B200 box: per-variant attention error structure + kernel timings.

Usage (under gpurun):  python tools/gpu_diag.py [attn|time|all]
Each section runs in its own subprocess with a timeout so one trap does not hide the rest.
Numbers printed here are development aids, never bench values.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _attn_case(variant, B, H, Nq, Nk, pattern):
    import torch

    from mvoc_b200 import ops
    from oracle import ops_ref

    torch.manual_seed(0)
    C = H * 64
    dev = torch.device("cuda:0")
    if pattern == "rand":
        q = torch.randn(B, Nq, C)
        k = torch.randn(B, Nk, C)
        v = torch.randn(B, Nk, C)
    elif pattern == "vident":  # uniform softmax, V = one-hot-ish ramps: isolates the P.V layout
        q = torch.zeros(B, Nq, C)
        k = torch.zeros(B, Nk, C)
        v = torch.randn(B, Nk, C)
    q, k, v = q.bfloat16(), k.bfloat16(), v.bfloat16()
    ref = ops_ref.sdpa_ref(q.float(), k.float(), v.float(), H)
    out = ops.attention(q.to(dev), k.to(dev), v.to(dev), H, variant=variant)
    torch.cuda.synchronize()
    o = out.float().cpu()
    err = (o - ref).norm() / ref.norm()
    print(f"variant={variant} B={B} H={H} Nq={Nq} Nk={Nk} {pattern}: rel_l2={err:.4e} finite={bool(torch.isfinite(o).all())}")
    if err > 1e-2:
        d = (o - ref).abs()
        print("  per-head err:", [f"{float(d[..., hh*64:(hh+1)*64].mean()):.3e}" for hh in range(H)])
        rows = d[0].mean(-1)
        print("  row-block(32) err:", [f"{float(rows[i:i+32].mean()):.2e}" for i in range(0, min(Nq, 256), 32)])
        cols = d[0, :, :64].mean(0)
        print("  col(8) err:", [f"{float(cols[i:i+8].mean()):.2e}" for i in range(0, 64, 8)])
        print("  out[0,0,:8]", o[0, 0, :8].tolist())
        print("  ref[0,0,:8]", ref[0, 0, :8].tolist())


def section_attn():
    for variant in (1, 2):
        for (B, H, Nq, Nk, pat) in [
            (1, 1, 128, 128, "vident"),
            (1, 1, 128, 128, "rand"),
            (1, 2, 256, 384, "rand"),
            (1, 1, 128, 145, "rand"),
            (2, 5, 1024, 1024, "rand"),
        ]:
            code = (
                "import sys; sys.path.insert(0, %r); from tools.gpu_diag import _attn_case; "
                "_attn_case(%d,%d,%d,%d,%d,%r)" % (ROOT, variant, B, H, Nq, Nk, pat)
            )
            try:
                r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
                print(r.stdout.strip())
                if r.returncode != 0:
                    print(f"  [exit {r.returncode}] " + r.stderr.strip()[-1500:])
            except subprocess.TimeoutExpired:
                print(f"variant={variant} {B},{H},{Nq},{Nk},{pat}: TIMEOUT")
            sys.stdout.flush()


def _time_cuda(fn, iters=20, warmup=3):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def section_time():
    import torch
    import torch.nn.functional as F

    from mvoc_b200 import ops

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    print("== attention (ms, TFLOP/s) vs torch SDPA ==")
    for (B, H, N, Nk) in [(80, 5, 4096, 4096), (80, 10, 1024, 1024), (80, 20, 256, 256), (80, 5, 4096, 145)]:
        C = H * 64
        q = torch.randn(B, N, C, device=dev).bfloat16()
        k = torch.randn(B, Nk, C, device=dev).bfloat16()
        v = torch.randn(B, Nk, C, device=dev).bfloat16()
        fl = 4.0 * B * H * N * Nk * 64
        for variant in (1, 2, 3, 4, 5):
            try:
                t = _time_cuda(lambda: ops.attention(q, k, v, H, variant=variant), iters=5)
                print(f"  mvoc v{variant} B={B} H={H} N={N} Nk={Nk}: {t:.3f} ms  {fl / t / 1e9:.1f} TF/s")
            except Exception as ex:  # noqa
                print(f"  mvoc v{variant} failed: {ex}")
        qh = q.view(B, N, H, 64).transpose(1, 2)
        kh = k.view(B, Nk, H, 64).transpose(1, 2)
        vh = v.view(B, Nk, H, 64).transpose(1, 2)
        t = _time_cuda(lambda: F.scaled_dot_product_attention(qh, kh, vh), iters=5)
        print(f"  torch SDPA: {t:.3f} ms  {fl / t / 1e9:.1f} TF/s")
    print("== temporal attention (ms, GB/s) ==")
    for (P, T, H) in [(20480, 16, 5), (5120, 16, 10)]:
        C = H * 64
        q = torch.randn(P, T, C, device=dev).bfloat16()
        k, v = torch.randn_like(q), torch.randn_like(q)
        by = 4.0 * P * T * C * 2
        t = _time_cuda(lambda: ops.temporal_attention(q, k, v, H))
        print(f"  mvoc P={P} T={T} H={H}: {t:.4f} ms  {by / t / 1e6:.0f} GB/s")
        qh = q.view(P, T, H, 64).transpose(1, 2)
        kh = k.view(P, T, H, 64).transpose(1, 2)
        vh = v.view(P, T, H, 64).transpose(1, 2)
        t = _time_cuda(lambda: F.scaled_dot_product_attention(qh, kh, vh))
        print(f"  torch SDPA: {t:.4f} ms  {by / t / 1e6:.0f} GB/s")
    print("== GroupNorm+SiLU (ms, GB/s) ==")
    for (N, C, H, W, fps) in [(80, 320, 64, 64, 1), (80, 960, 64, 64, 1), (80, 320, 64, 64, 16), (80, 640, 32, 32, 1)]:
        x = torch.randn(N, C, H, W, device=dev).bfloat16()
        w = torch.ones(C, device=dev).bfloat16()
        b = torch.zeros(C, device=dev).bfloat16()
        by = 2.0 * x.numel() * 2
        t = _time_cuda(lambda: ops.groupnorm_silu(x, w, b, 32, 1e-5, True, fps))
        print(f"  mvoc N={N} C={C} {H}x{W} fps={fps}: {t:.4f} ms  {by / t / 1e6:.0f} GB/s")
        if fps == 1:
            t = _time_cuda(lambda: F.silu(F.group_norm(x, 32, w, b, 1e-5)))
            print(f"  torch GN+SiLU: {t:.4f} ms  {by / t / 1e6:.0f} GB/s")
    print("== channels-last GroupNorm / LayerNorm / GEGLU (ms, GB/s of algorithmic bytes) ==")
    for (N, C, H, W, fps) in [(80, 320, 64, 64, 1), (80, 960, 64, 64, 1), (80, 320, 64, 64, 16), (80, 640, 32, 32, 1)]:
        x = torch.randn(N, H, W, C, device=dev).bfloat16()
        w = torch.ones(C, device=dev).bfloat16()
        b = torch.zeros(C, device=dev).bfloat16()
        by = 2.0 * x.numel() * 2
        t = _time_cuda(lambda: ops.groupnorm_nhwc(x, w, b, 32, 1e-5, True, fps))
        print(f"  gn_nhwc N={N} C={C} {H}x{W} fps={fps}: {t:.4f} ms  {by / t / 1e6:.0f} GB/s")
    for (M, C) in [(327680, 320), (81920, 640), (20480, 1280)]:
        x = torch.randn(M, C, device=dev).bfloat16()
        w = torch.ones(C, device=dev).bfloat16()
        b = torch.zeros(C, device=dev).bfloat16()
        by = 2.0 * x.numel() * 2
        t = _time_cuda(lambda: ops.layernorm(x, w, b, 1e-5))
        t2 = _time_cuda(lambda: F.layer_norm(x, (C,), w, b, 1e-5))
        print(f"  layernorm M={M} C={C}: {t:.4f} ms  {by / t / 1e6:.0f} GB/s   (torch {t2:.4f} ms)")
    x = torch.randn(327680, 2560, device=dev).bfloat16()
    t = _time_cuda(lambda: ops.geglu(x))
    print(f"  geglu M=327680 F=1280: {t:.4f} ms  {3.0 * 327680 * 1280 * 2 / t / 1e6:.0f} GB/s")
    del x
    print("== blends ==")
    nb, T, hw, C = 5, 16, 4096, 320
    q = torch.randn(nb * T, hw, C, device=dev).bfloat16()
    k = torch.randn_like(q)
    m = (torch.rand(2, T * hw, device=dev) < 0.1).to(torch.uint8)
    by = 2 * (2 * T * hw * C * 2 + T * hw * C * 2)  # read 1 + write 2 copies, for Q and K
    t = _time_cuda(lambda: ops.qk_blend_(q, k, m, 2, False))
    print(f"  qk_blend spatial l0: {t:.4f} ms  {by / t / 1e6:.0f} GB/s (select path bytes)")
    x = torch.randn(nb * T, 320, 64, 64, device=dev).bfloat16()
    m3 = (torch.rand(2, T, 4096, device=dev) < 0.1).to(torch.uint8)
    by = 3.0 * T * 320 * 4096 * 2
    t = _time_cuda(lambda: ops.feature_blend_(x, m3, 2, T))
    print(f"  feature_blend l0: {t:.4f} ms  {by / t / 1e6:.0f} GB/s (min bytes)")


def section_pair():
    """Injected l0 / l1 spatial self-attention: all 5 branches through the plain kernel vs sources plain + the
    uncond / cond pair through the one-softmax pair kernel (what mvoc_attn_inject_fwd does with share_p)."""
    import torch

    from mvoc_b200 import ops

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for (T, H, N) in [(16, 5, 4096), (16, 10, 1024), (16, 20, 256)]:
        C = H * 64
        nb = 5
        q = torch.randn(nb * T, N, C, device=dev).bfloat16()
        k = torch.randn(nb * T, N, C, device=dev).bfloat16()
        v = torch.randn(nb * T, N, C, device=dev).bfloat16()
        fl = 4.0 * nb * T * H * N * N * 64
        t_all = _time_cuda(lambda: ops.attention(q, k, v, H), iters=5)
        t_src = _time_cuda(lambda: ops.attention(q[:3 * T], k[:3 * T], v[:3 * T], H), iters=5)
        t_pair = _time_cuda(lambda: ops.attention_pair(q[3 * T:4 * T], k[3 * T:4 * T], v[3 * T:], H, T), iters=5)
        t_two = _time_cuda(lambda: ops.attention(q[3 * T:], k[3 * T:], v[3 * T:], H), iters=5)
        print(f"  T={T} H={H} N={N}: all-plain {t_all:.3f} ms ({fl / t_all / 1e9:.0f} TF/s) | sources {t_src:.3f} + pair "
              f"{t_pair:.3f} = {t_src + t_pair:.3f} ms ({fl / (t_src + t_pair) / 1e9:.0f} TF/s of reference FLOPs); "
              f"the pair as two plain branches: {t_two:.3f} ms")


def section_dense():
    """The tcgen05 GEMM family (csrc/gemm_tc.cu) against the libraries it replaces, on the UNet's shapes."""
    import torch
    import torch.nn.functional as F

    from mvoc_b200 import ops

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    variants = [int(v) for v in os.environ.get("MVOC_DIAG_VARIANTS", "0,1").split(",")]
    print("== 3x3 conv, channels-last (ms, TFLOP/s): tcgen05 implicit GEMM vs cuDNN ==")
    for (N, H, W, ci, co) in [(80, 64, 64, 320, 320), (80, 64, 64, 640, 320), (80, 64, 64, 960, 320),
                              (80, 32, 32, 640, 640), (80, 32, 32, 1280, 640), (80, 32, 32, 1920, 640),
                              (80, 16, 16, 1280, 1280), (80, 16, 16, 2560, 1280), (80, 8, 8, 1280, 1280),
                              (80, 8, 8, 2560, 1280)]:
        x = torch.randn(N, H, W, ci, device=dev).bfloat16()
        w = (torch.randn(co, ci, 3, 3, device=dev) * (9 * ci) ** -0.5).bfloat16()
        b = torch.randn(co, device=dev).bfloat16()
        wt = ops.conv_taps(w)
        wcl = w.contiguous(memory_format=torch.channels_last)
        fl = 2.0 * N * H * W * co * ci * 9
        t = _time_cuda(lambda: F.conv2d(x.permute(0, 3, 1, 2), wcl, b, padding=1), iters=5)
        print(f"  cuDNN        {N}x{H}x{W} {ci}->{co}: {t:.3f} ms  {fl / t / 1e9:.1f} TF/s")
        for v in variants:
            t = _time_cuda(lambda: ops.conv3x3(x, wt, b, variant=v), iters=5)
            print(f"  tcgen05 v{v}   {N}x{H}x{W} {ci}->{co}: {t:.3f} ms  {fl / t / 1e9:.1f} TF/s")
    print("== Linear (ms, TFLOP/s): mvoc_linear (bias [+ residual] in the epilogue) vs cuBLAS (+ add_) ==")
    for (M, K, N, res) in [(327680, 320, 960, 0), (327680, 320, 320, 1), (327680, 1280, 320, 1), (81920, 640, 1920, 0),
                           (81920, 640, 640, 1), (81920, 2560, 640, 1), (20480, 1280, 3840, 0), (20480, 1280, 1280, 1),
                           (20480, 5120, 1280, 1), (11600, 1024, 640, 0), (5120, 1280, 1280, 1)]:
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        r = torch.randn(M, N, device=dev).bfloat16() if res else None
        fl = 2.0 * M * K * N
        t = _time_cuda(lambda: F.linear(x, w, b).add_(r) if res else F.linear(x, w, b), iters=5)
        print(f"  cuBLAS       M={M} K={K} N={N}{' +res' if res else ''}: {t:.3f} ms  {fl / t / 1e9:.1f} TF/s")
        for v in variants:
            t = _time_cuda(lambda: ops.linear(x, w, b, r, variant=v), iters=5)
            print(f"  tcgen05 v{v}   M={M} K={K} N={N}{' +res' if res else ''}: {t:.3f} ms  {fl / t / 1e9:.1f} TF/s")
    print("== GEGLU projection (ms): gate in the GEMM epilogue vs cuBLAS + mvoc_geglu ==")
    for (M, K) in [(327680, 320), (81920, 640), (20480, 1280)]:
        Fo = 4 * K
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(2 * Fo, K, device=dev) * K ** -0.5).bfloat16()
        b = torch.randn(2 * Fo, device=dev).bfloat16()
        fl = 2.0 * M * K * 2 * Fo
        t0 = _time_cuda(lambda: ops.geglu(F.linear(x, w, b)), iters=5)
        print(f"  cuBLAS+geglu M={M} K={K} F={Fo}: {t0:.3f} ms  {fl / t0 / 1e9:.1f} TF/s")
        for v in variants:
            t1 = _time_cuda(lambda: ops.linear_geglu(x, w, b, variant=v), iters=5)
            print(f"  tcgen05 v{v}   M={M} K={K} F={Fo}: {t1:.3f} ms  {fl / t1 / 1e9:.1f} TF/s")
    print("== temporal conv (3,1,1) (ms): one kernel vs 3 cuBLAS addmm per video ==")
    for (B, T, S, C) in [(5, 16, 4096, 320), (5, 16, 1024, 640), (5, 16, 256, 1280), (5, 16, 64, 1280)]:
        x = torch.randn(B * T, S, C, device=dev).bfloat16()
        w = (torch.randn(C, C, 3, 1, 1, device=dev) * (3 * C) ** -0.5).bfloat16()
        b = torch.randn(C, device=dev).bfloat16()
        wt = ops.conv_taps(w)
        fl = 2.0 * B * T * S * C * C * 3
        for v in variants:
            t = _time_cuda(lambda: ops.temporal_conv3(x, wt, b, B, T, variant=v), iters=5)
            print(f"  tcgen05 v{v}   B={B} T={T} S={S} C={C}: {t:.3f} ms  {fl / t / 1e9:.1f} TF/s")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "dense":
        section_dense()
    if what == "pair":
        section_pair()
    if what in ("attn", "all"):
        section_attn()
    if what in ("time", "all"):
        section_time()
