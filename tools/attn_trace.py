"""Phase timing of the attention kernel from its own clock64() trace (mvoc_attn_fwd_trace): where one CTA spends
the cycles of a key block.  Development aid; run under gpurun."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mvoc_b200 import _cabi  # noqa: E402


def main():
    B, H, N = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (80, 5, 4096)
    pair = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    C = H * 64
    q = torch.randn(B, N, C, device=dev).bfloat16()
    k = torch.randn(B, N, C, device=dev).bfloat16()
    v = torch.randn(B + pair, N, C, device=dev).bfloat16()
    o = torch.empty(B + pair, N, C, device=dev).bfloat16()
    bn = 64 if pair else 128
    nblk = (N + bn - 1) // bn
    trace = torch.zeros(nblk * 8, dtype=torch.int64, device=dev)
    st = (ctypes.c_int64 * 12)(*([N * C, C, 64] * 4))
    lib = _cabi.load()
    for _ in range(3):
        rc = lib.mvoc_attn_fwd_trace(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, N, 64, st, pair,
                                     0.125, 0, 0, trace.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "mvoc_attn_fwd_trace")
    torch.cuda.synchronize()
    t = trace.view(nblk, 8).cpu()
    names = ["wait S", "load+max", "exp loop", "wait PV", "store P"]
    print(f"B={B} H={H} N={N} pair={pair}: clocks of CTA (0,0,0), softmax thread 0, per key block ({bn} keys)")
    tot = [0.0] * 6
    cnt = 0
    for j in range(1, nblk - 1):
        e = t[j]
        d = [int(e[1] - e[0]), int(e[2] - e[1]), int(e[3] - e[2]), int(e[4] - e[3]), int(e[5] - e[4])]
        period = int(t[j + 1][0] - e[0])
        qk_lead = int(e[1] - e[6]) if j + 0 < nblk else 0      # S(j) ready minus QK(j) issue
        if j < 4 or j % 8 == 0:
            print(f"  block {j:3d}: " + "  ".join(f"{n} {x:5d}" for n, x in zip(names, d)) +
                  f" | period {period:5d} | QK issue -> S ready {qk_lead:5d} | PV issue after P {int(e[7] - e[5]):5d}")
        for i, x in enumerate(d):
            tot[i] += x
        tot[5] += period
        cnt += 1
    print("  mean:      " + "  ".join(f"{n} {tot[i] / cnt:7.0f}" for i, n in enumerate(names)) +
          f" | period {tot[5] / cnt:7.0f}")


if __name__ == "__main__":
    main()
