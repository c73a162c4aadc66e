"""Launches ONE kernel family of the library a few times at its benchmark shape (the target of an
`ncu --set full -k regex:<kernel> -s 2 -c 1` capture).  usage: python tools/ncu_kernels.py <name>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mvoc_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
name = sys.argv[1]
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
REPS = 4


def run(fn):
    for _ in range(REPS):
        fn()
    torch.cuda.synchronize()


if name == "attn":
    qkv = bf(80, 4096, 960)
    run(lambda: ops.attention(qkv[..., :320], qkv[..., 320:640], qkv[..., 640:], 5))
elif name == "attn_pair":
    q, k, v = bf(16, 4096, 320), bf(16, 4096, 320), bf(32, 4096, 320)
    run(lambda: ops.attention_pair(q, k, v, 5, 16))
elif name == "attn_temporal":
    q, k, v = bf(5 * 16 * 4096, 320), bf(5 * 16 * 4096, 320), bf(5 * 16 * 4096, 320)
    run(lambda: ops.temporal_attention_frames(q, k, v, 5, 5, 16, 4096))
elif name in ("gn", "gn_t"):
    x, w, b = bf(80, 64, 64, 320), bf(320), bf(320)
    run(lambda: ops.groupnorm_nhwc(x, w, b, 32, 1e-5, True, 16 if name == "gn_t" else 1))
elif name == "layernorm":
    x, w, b = bf(327680, 320), bf(320), bf(320)
    run(lambda: ops.layernorm(x, w, b, 1e-5))
elif name == "qk_blend":
    q, k = bf(80, 4096, 320), bf(80, 4096, 320)
    m = (torch.rand(2, 16 * 4096, device=dev) < 0.1).to(torch.uint8)
    run(lambda: ops.qk_blend_(q, k, m, 2, False))
elif name == "feature_blend":
    x = bf(80, 320, 64, 64)
    m = (torch.rand(2, 16, 4096, device=dev) < 0.1).to(torch.uint8)
    run(lambda: ops.feature_blend_(x, m, 2, 16))
elif name == "conv":
    x, w, b = bf(80, 64, 64, 320), ops.conv_taps(bf(320, 320, 3, 3) * 0.02), bf(320)
    run(lambda: ops.conv3x3(x, w, b))
elif name == "conv_l2":
    x, w, b = bf(80, 16, 16, 1280), ops.conv_taps(bf(1280, 1280, 3, 3) * 0.01), bf(1280)
    run(lambda: ops.conv3x3(x, w, b))
elif name == "linear":
    x, w, b, r = bf(327680, 1280), bf(320, 1280) * 0.03, bf(320), bf(327680, 320)
    run(lambda: ops.linear(x, w, b, r))
elif name == "geglu":
    x, w, b = bf(327680, 320), bf(2560, 320) * 0.05, bf(2560)
    run(lambda: ops.linear_geglu(x, w, b))
elif name == "tconv":
    x, w, b = bf(80, 4096, 320), ops.conv_taps(bf(320, 320, 3, 1, 1) * 0.03), bf(320)
    run(lambda: ops.temporal_conv3(x, w, b, 5, 16))
else:
    raise SystemExit(f"unknown kernel family {name}")
print("done", name)
