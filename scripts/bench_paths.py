#!/usr/bin/env python
"""Throughput of the other hot-path rows (BASELINE.json configs[2..4] shapes) on one B200:
PolicyNetwork1UNet logprob fwd+bwd (b=25), PolicyNetwork2UNet IL logits fwd+bwd (b=20) and critic
fwd+bwd, ResnetFeatureExtractor forward (25 frames of 224x224), EncoderBlock fwd+bwd
(E=3072, 256 tokens). CUDA-event timing, 3 warm-up + N timed iterations; prints one JSON line per
path with the algorithmic FLOP figures of SURVEY.md §8d."""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reinformcement-optimized-video-reconstruction_b200"))
import torch  # noqa: E402

warnings.filterwarnings("ignore")
dev = torch.device("cuda:0")
import _native  # noqa: E402
_native.require_device()


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _native.lib.rovr_launch_count()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (_native.lib.rovr_launch_count() - n0) / iters


def report(name, ms, launches, units, unit, gflop_per_unit, fn=None, inputs=None, modules=()):
    rate = units / (ms * 1e-3)
    line = {"path": name, "ms": round(ms, 3), "launches": launches, "rate": round(rate, 1), "unit": unit,
            "tflops": round(rate * gflop_per_unit / 1e3, 2)}
    if fn is not None:      # the same step captured into one CUDA graph (graphs.GraphedFunction)
        from graphs import GraphedFunction
        g = GraphedFunction(fn, inputs, modules=modules)
        gms, _ = timeit(lambda: g(*inputs))
        line.update({"graph_ms": round(gms, 3), "graph_rate": round(units / (gms * 1e-3), 1),
                     "graph_tflops": round(units / (gms * 1e-3) * gflop_per_unit / 1e3, 2)})
    print(json.dumps(line), flush=True)


def main():
    torch.manual_seed(0)
    from policy_net_1 import PolicyNetwork1UNet
    from policy_net_2 import PolicyNetwork2UNet
    from resnet_extractor import ResnetFeatureExtractor
    from common_layers import EncoderBlock

    pn1 = PolicyNetwork1UNet().to(dev).train()
    img, ctx = torch.rand(25, 3, 80, 80, device=dev), torch.rand(25, 3, 80, 80, device=dev)
    act = torch.randint(0, 25, (25,), device=dev)

    def f_pn1():
        pn1.zero_grad(set_to_none=True)
        pn1.logprob(img, ctx, act).sum().backward()
    ms, l = timeit(f_pn1)

    def g_pn1(i, c, a):
        pn1.logprob(i, c, a).sum().backward()
    pn1.zero_grad(set_to_none=True)
    report("PN1 logprob fwd+bwd b=25 80x80", ms, l, 25, "samples/s", 2.960, g_pn1, (img, ctx, act), [pn1])

    pn2 = PolicyNetwork2UNet().to(dev).train()
    enc, feat = torch.rand(20, 1, 160, 160, device=dev), torch.randn(20, 1, 1024, device=dev)
    tgt = torch.arange(20, device=dev).view(20, 1, 1)

    def f_pn2():
        pn2.zero_grad(set_to_none=True)
        (pn2(enc, feat, tgt, extra=True) ** 2).sum().backward()
    ms, l = timeit(f_pn2)

    def g_pn2(e, f, t):
        (pn2(e, f, t, extra=True) ** 2).sum().backward()
    pn2.zero_grad(set_to_none=True)
    report("PN2 IL logits fwd+bwd b=20 160x160", ms, l, 20, "samples/s", 0.503, g_pn2, (enc, feat, tgt), [pn2])
    for k in (8, 64):
        b = 20 * k
        encb, featb = torch.rand(b, 1, 160, 160, device=dev), torch.randn(b, 1, 1024, device=dev)
        crit = PolicyNetwork2UNet(is_critic=True).to(dev).train()

        def f_crit():
            crit.zero_grad(set_to_none=True)
            (crit(encb[:, 0], featb[:, 0], None) ** 2).sum().backward()
        ms, l = timeit(f_crit, iters=10)
        report(f"PN2 critic fwd+bwd b={b}", ms, l, b, "samples/s", 0.503)

    rn = ResnetFeatureExtractor(pretrained=False).to(dev)
    rn.resnet.eval()
    for p in rn.resnet.parameters():
        p.requires_grad = False
    frames = torch.rand(1, 25, 3, 224, 224, device=dev)

    def f_rn():
        rn.zero_grad(set_to_none=True)
        (rn(frames) ** 2).sum().backward()
    ms, l = timeit(f_rn, iters=10)

    def g_rn(fr):
        (rn(fr) ** 2).sum().backward()
    rn.zero_grad(set_to_none=True)
    report("ResNet-50 extractor fwd (+linear bwd) 25 frames 224x224", ms, l, 25, "frames/s", 8.174, g_rn, (frames,), [rn])

    E, S = 3072, 256
    for B in (8, 24):
        blk = EncoderBlock(E, 8, 0.0).to(dev).eval()
        x = torch.randn(B, S, E, device=dev, requires_grad=True)

        def f_enc():
            blk.zero_grad(set_to_none=True)
            (blk(x) ** 2).sum().backward()
        ms, l = timeit(f_enc, iters=10)

        def g_enc(xx):
            (blk(xx) ** 2).sum().backward()
        blk.zero_grad(set_to_none=True)
        x.grad = None
        report(f"EncoderBlock fwd+bwd E=3072 S=256 B={B} heads=8", ms, l, B, "sequences/s", 3 * (20.13 + 2.42),
               g_enc, (x,), [blk])
        del blk, x

if __name__ == "__main__":
    main()
