"""Developer bench of the VAE (row N4): decode / encode time on the GPU (CUDA events) and the CPU oracle beside it.
Usage (GPU box): python tools/vae_bench.py [--size 512] [--batch 1] [--cpu]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def conv_flops(vae, H, W, decode):
    """2 * MACs of every convolution / linear / attention GEMM of one image (algorithmic: real channel counts)."""
    cfg = vae.config
    ch = list(cfg.block_out_channels)
    total = 0

    def conv(cin, cout, h, w, k=9):
        nonlocal total
        total += 2 * k * cin * cout * h * w

    def resnet(cin, cout, h, w):
        conv(cin, cout, h, w)
        conv(cout, cout, h, w)
        if cin != cout:
            conv(cin, cout, h, w, 1)

    def mid(c, h, w):
        resnet(c, c, h, w)
        resnet(c, c, h, w)
        s = h * w
        nonlocal total
        total += 2 * 4 * c * c * s + 2 * 2 * s * s * c

    if decode:
        rev = ch[::-1]
        h, w = H, W
        conv(cfg.latent_channels, rev[0], h, w)
        mid(rev[0], h, w)
        cin = rev[0]
        for i, cout in enumerate(rev):
            for j in range(cfg.layers_per_block + 1):
                resnet(cin if j == 0 else cout, cout, h, w)
            if i != len(rev) - 1:
                h, w = 2 * h, 2 * w
                conv(cout, cout, h, w)
            cin = cout
        conv(rev[-1], cfg.out_channels, h, w)
    else:
        h, w = H, W
        conv(cfg.in_channels, ch[0], h, w)
        cin = ch[0]
        for i, cout in enumerate(ch):
            for j in range(cfg.layers_per_block):
                resnet(cin if j == 0 else cout, cout, h, w)
            if i != len(ch) - 1:
                h, w = h // 2, w // 2
                conv(cout, cout, h, w)
            cin = cout
        mid(ch[-1], h, w)
        conv(ch[-1], 2 * cfg.latent_channels, h, w)
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle (B=1)")
    a = ap.parse_args()
    import dit4sr_b200
    from dit4sr_b200 import ops
    from oracle import vae_oracle as VO
    cfg = VO.sd3_vae_config()
    sd = VO.init_state_dict(cfg, seed=0)
    vae = dit4sr_b200.AutoencoderKL()
    vae.load_state_dict(sd)
    vae = vae.to("cuda").eval()
    z = torch.randn(a.batch, 16, a.size // 8, a.size // 8, device="cuda")
    img = torch.rand(a.batch, 3, a.size, a.size, device="cuda") * 2 - 1
    for name, fn, fl in (("decode", lambda: vae.decode(z), conv_flops(vae, a.size // 8, a.size // 8, True)),
                         ("encode", lambda: vae.encode(img), conv_flops(vae, a.size, a.size, False))):
        fn()
        torch.cuda.synchronize()
        n0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        print(f"vae {name} {a.size}px batch {a.batch}: {ms:.3f} ms ({a.batch / ms * 1e3:.1f} images/s), "
              f"{fl * a.batch / 1e12:.3f} TFLOP -> {fl * a.batch / ms / 1e9:.0f} TFLOP/s, "
              f"{(ops.launch_count() - n0) // a.reps} launches, peak memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
    if a.cpu:
        torch.set_num_threads(os.cpu_count())
        zc = z[:1].cpu()
        with torch.no_grad():
            VO.decoder_forward(sd, cfg, zc)
            t0 = time.perf_counter()
            VO.decoder_forward(sd, cfg, zc)
            t1 = time.perf_counter()
        print(f"cpu oracle decode {a.size}px batch 1 ({os.cpu_count()} threads): {(t1 - t0) * 1e3:.0f} ms")


if __name__ == "__main__":
    main()
