"""Runs the LocalNet layers one by one at the BASELINE size (B=24, 256x256) so that ncu can
capture each tcgen05 launch in isolation: python scripts/prof_layers.py [layer ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reinformcement-optimized-video-reconstruction_b200"))
import torch  # noqa: E402
import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
B = 24
want = set(sys.argv[1:])


def act(h, c):
    return (torch.randn((B, h, h, c), device=dev) * 0.5).clamp_min(0).to(BF)


def conv_layer(name, h, cin, cout):
    x, dy = act(h, cin), act(h, cout)
    w = torch.randn((cout, cin, 3, 3), device=dev) * 0.05
    b = torch.zeros(cout, device=dev)
    wk, wd = ops.repack_conv3x3(w), ops.repack_conv3x3(w, True)
    y, dx, dw = torch.empty_like(dy), torch.empty_like(x), torch.empty_like(w)
    for _ in range(2):
        ops.conv3x3_fprop(x, wk, b, y)
        ops.conv3x3_dgrad(dy, wd, dx, mask=x)
        ops.conv3x3_wgrad(dy, x, dw)
    torch.cuda.synchronize()


def up_layer(name, h, cin, cout):
    x, dy = act(h, cin), act(2 * h, cout)
    w = torch.randn((cin, cout, 2, 2), device=dev) * 0.05
    b = torch.zeros(cout, device=dev)
    wk, wd = ops.repack_convT2x2(w), ops.repack_convT2x2(w, True)
    y, dx, dw = torch.empty_like(dy), torch.empty_like(x), torch.empty_like(w)
    for _ in range(2):
        ops.convT2x2_fprop(x, wk, b, y)
        ops.convT2x2_dgrad(dy, wd, dx, mask=x)
        ops.convT2x2_wgrad(dy, x, dw)
    torch.cuda.synchronize()


LAYERS = {"conv1": (conv_layer, 256, 16, 64), "conv2": (conv_layer, 128, 64, 128),
          "conv3": (conv_layer, 64, 128, 256), "conv4": (conv_layer, 32, 256, 512),
          "upconv1": (up_layer, 32, 512, 256), "conv5": (conv_layer, 64, 512, 256),
          "upconv2": (up_layer, 64, 256, 128), "conv6": (conv_layer, 128, 256, 128),
          "upconv3": (up_layer, 128, 128, 64), "conv7": (conv_layer, 256, 128, 64)}
for name, (fn, h, cin, cout) in LAYERS.items():
    if not want or name in want:
        fn(name, h, cin, cout)
print("done")
