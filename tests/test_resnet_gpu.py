"""GPU parity of the ResnetFeatureExtractor drop-in against fixtures produced by the unmodified
reference class (tests/golden/make_golden.py::gen_resnet: seeded weights, eval + frozen trunk as
with pretrained=True) and of its helper kernels against PyTorch.

Tolerance: bf16 tensor-core trunk (53 convolutions) -> 2e-2 L2-relative on the feature map and on
the gradients of `linear`; the GPU resampler must reproduce PIL's 8-bit output exactly.
"""
import os
import warnings

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import rovr_oracle as O

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _dev():
    import _native
    _native.require_device()
    return torch.device("cuda:0")


def _rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).norm() / (ref.norm() + 1e-20)).item()


def _inputs():
    g = torch.Generator().manual_seed(62)
    return {"a": torch.rand((1, 3, 3, 48, 64), generator=g), "b": torch.rand((1, 2, 3, 224, 224), generator=g),
            "c": torch.rand((1, 1, 3, 256, 256), generator=g)}


def _module(dev):
    from resnet_extractor import ResnetFeatureExtractor
    warnings.filterwarnings("ignore")
    torch.manual_seed(0)
    m = ResnetFeatureExtractor(pretrained=False)
    O.resnet_randomise_bn(m.resnet, 61)
    m.resnet.eval()
    for p in m.resnet.parameters():
        p.requires_grad = False
    return m.to(dev)


def test_resnet_helpers():
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, 14, 18, 64), generator=g).to(dev).to(BF)
    xn = x.float().permute(0, 3, 1, 2)
    y = ops.maxpool_pad_fwd(x, 3, 2, 1)
    assert torch.equal(y.float().permute(0, 3, 1, 2), F.max_pool2d(xn, 3, 2, 1))
    s = ops.subsample(x, 2)
    assert torch.equal(s, x[:, ::2, ::2].contiguous())
    b = torch.randn((2, 14, 18, 64), generator=g).to(dev).to(BF)
    r = ops.add_relu(x, b)
    assert torch.equal(r, F.relu(x.float() + b.float()).to(BF))
    a = ops.avgpool(x)
    assert _rel(a, xn.mean(dim=(2, 3))) < 1e-5
    # BN folding
    w = torch.randn((32, 16, 3, 3), generator=g).to(dev)
    gam, bet = torch.rand(32, generator=g).to(dev) + 0.5, torch.randn(32, generator=g).to(dev)
    mu, var = torch.randn(32, generator=g).to(dev), torch.rand(32, generator=g).to(dev) + 0.5
    wf, bf = ops.fold_bn(w, gam, bet, mu, var, 1e-5)
    sc = gam / torch.sqrt(var + 1e-5)
    assert _rel(wf, w * sc[:, None, None, None]) < 1e-6 and _rel(bf, bet - mu * sc) < 1e-6
    # stem im2col + GEMM == 7x7 stride-2 pad-3 convolution
    img = torch.rand((2, 3, 40, 56), generator=g).to(dev)
    w7 = (torch.randn((64, 3, 7, 7), generator=g) * 0.1).to(dev)
    cols, Ho, Wo = ops.stem_im2col(img, 160, quantise=False)
    wk = ops.repack_linear(w7.reshape(64, -1), False)
    out = ops.gemm_bf16(cols, wk, None, relu=False).view(2, Ho, Wo, 64)
    ref = F.conv2d(img.to(BF).float(), w7.to(BF).float(), stride=2, padding=3)
    assert _rel(out.float().permute(0, 3, 1, 2), ref) < 1e-2
    # mosaic paste / gather
    feat = torch.randn((6, 768), generator=g).to(dev)
    fmap = torch.zeros((2, 3, 80, 80), device=dev)
    ops.mosaic_paste(feat, fmap, slots_per_mosaic=3)
    for r_ in range(6):
        bi, s_ = r_ // 3, r_ % 3
        assert torch.equal(fmap[bi, :, s_ // 5 * 16:s_ // 5 * 16 + 16, s_ % 5 * 16:s_ % 5 * 16 + 16], feat[r_].view(3, 16, 16))
    back = torch.empty_like(feat)
    ops.mosaic_paste(back, fmap, slots_per_mosaic=3, gather=True)
    assert torch.equal(back, feat)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_resize_matches_pil(golden_dir, tag):
    """ToPILImage -> Resize((224,224)) -> ToTensor on the GPU == the reference's host PIL path."""
    import ops
    dev = _dev()
    G = np.load(os.path.join(golden_dir, "resnet.npz"), allow_pickle=False)
    x = _inputs()[tag][0, :1].to(dev)
    y = ops.resize_antialias(x.contiguous(), 224, 224)[0].cpu().numpy()[:, ::7, ::5]
    ref = G[f"{tag}/prep0"]
    diff = np.abs(y - ref) * 255
    print(f"resize[{tag}]: max diff {diff.max():.3f} levels, mismatching {float((diff > 0.5).mean()):.4%}")
    assert diff.max() < 0.5, "GPU resampler differs from PIL's 8-bit output"


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_resnet_extractor_vs_reference_golden(golden_dir, tag):
    dev = _dev()
    G = np.load(os.path.join(golden_dir, "resnet.npz"), allow_pickle=False)
    m = _module(dev)
    assert list(m.state_dict().keys()) == list(G["keys"])
    x = _inputs()[tag].to(dev)
    y = m(x)
    (y ** 2).sum().backward()
    ref = torch.from_numpy(G[f"{tag}/y"])
    r = _rel(y, ref)
    print(f"resnet[{tag}] feature map l2-rel {r:.3e}")
    assert y.shape == (1, 3, 80, 80) and r < 2e-2
    S = x.shape[1]
    used = torch.zeros(80, 80, dtype=torch.bool)
    for s in range(S):
        used[s // 5 * 16:s // 5 * 16 + 16, s % 5 * 16:s % 5 * 16 + 16] = True
    assert float(y[0, :, ~used.to(dev)].abs().sum()) == 0.0, "pixels outside the pasted tiles must stay zero"
    for name in ("linear.weight", "linear.bias"):
        g = dict(m.named_parameters())[name].grad.cpu()
        key = f"{tag}/{name}"
        gn = float(G[f"gnorm/{key}"])
        assert abs(float(g.double().norm()) - gn) < 3e-2 * gn, name
        if f"gfull/{key}" in G:
            assert _rel(g, torch.from_numpy(G[f"gfull/{key}"])) < 3e-2, name
        else:
            idx = torch.from_numpy(G[f"gidx/{key}"])
            refv = torch.from_numpy(G[f"gval/{key}"])
            assert _rel(g.reshape(-1)[idx], refv) < 3e-2, name
    assert all(p.grad is None for p in m.resnet.parameters())


def test_resnet_encode_insert_extract(golden_dir):
    dev = _dev()
    m = _module(dev)
    x = _inputs()["b"].to(dev)
    fmap = m(x).detach()
    tile = m.encode(x[0, 1])
    assert tile.shape == (3, 16, 16) and _rel(tile, fmap[0, :, 0:16, 16:32]) < 1e-3
    enc = torch.zeros((1, 3, 80, 80), device=dev)
    enc = m.insert_encoded_frame_batch(torch.tensor([[7]]), x[:, 0], enc)
    assert _rel(enc[0, :, 16:32, 32:48], fmap[0, :, 0:16, 0:16]) < 1e-3
    patches = m.extract_patch([[0, 1]], fmap)
    assert patches.shape == (1, 2, 3, 16, 16) and torch.equal(patches[0, 1], fmap[0, :, 0:16, 16:32])
    m.resnet.train()                                    # frozen but train-mode trunk: batch statistics, forward-only
    assert torch.isfinite(m(x)).all() and _rel(m(x).detach(), fmap) > 1e-3
    with pytest.raises(RuntimeError):
        m.resnet.eval()
        m(x.cpu())


def test_resnet_built_like_the_reference(monkeypatch):
    """The ONLY way the reference builds the extractor is `ResnetFeatureExtractor(pretrained=True)`
    (rovr/rovr.py:31, rovr/imitation_learning.py:39): `.eval()` + frozen trunk inside the ctor, then
    a fresh nn.Sequential around the children (whose own `training` flag is True). No extra .eval()
    here. (There is no network for the weight download, so torchvision's factory is patched to
    ignore `pretrained`; the ctor path under test is unchanged.)"""
    import resnet_extractor as M
    real = M.models.resnet50
    monkeypatch.setattr(M.models, "resnet50", lambda pretrained=False: real())
    warnings.filterwarnings("ignore")
    torch.manual_seed(0)
    m = M.ResnetFeatureExtractor(pretrained=True).to(_dev())
    assert m.resnet.training and not any(c.training for c in m.resnet.children())   # the trap of ADVICE r1
    assert m.training                                   # the module itself was never put in eval mode
    x = torch.rand((1, 2, 3, 64, 64), generator=torch.Generator().manual_seed(1)).to(_dev())
    fmap = m(x)
    assert fmap.shape == (1, 3, 80, 80) and torch.isfinite(fmap).all()
    fmap.sum().backward()
    assert m.linear.weight.grad is not None and all(p.grad is None for p in m.resnet.parameters())
    tile = m.encode(x[0, 0])
    assert tile.shape == (3, 16, 16)
    # a default-constructed (train-mode, TRAINABLE) trunk is forward-only: with autograd on it says so ...
    m2 = M.ResnetFeatureExtractor().to(_dev())
    with pytest.raises(NotImplementedError):
        m2(x)
    # ... and under no_grad it runs with batch statistics (test_resnet_train_mode_trunk_vs_oracle checks the numbers)
    with torch.no_grad():
        assert torch.isfinite(m2(x)).all()


def test_resnet_train_mode_trunk_vs_oracle():
    """The default constructor (`pretrained=False`, rovr/resnet_extractor.py:6-8) leaves the trunk in TRAINING mode,
    and the reference encodes one frame per call (:42-47), so every BatchNorm normalises each frame with that frame's
    own statistics and updates its running buffers once per frame, in frame order. The drop-in batches the frames and
    must reproduce exactly that: feature map, running_mean / running_var and num_batches_tracked against the oracle
    (the reference's own loop over the same torchvision modules on the CPU)."""
    import copy
    import resnet_extractor as M
    dev = _dev()
    warnings.filterwarnings("ignore")
    torch.manual_seed(11)
    m = M.ResnetFeatureExtractor()
    O.resnet_randomise_bn(m.resnet, 29)                 # non-trivial affine and running statistics
    ref_seq = copy.deepcopy(m.resnet).train()
    lw, lb = m.linear.weight.detach().clone(), m.linear.bias.detach().clone()
    m = m.to(dev)
    x = torch.rand((2, 3, 3, 96, 128), generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        want = O.resnet_extractor_forward(ref_seq, lw, lb, x)
        got = m(x.to(dev))
    err = ((got.cpu() - want).norm() / want.norm()).item()
    print(f"train-mode trunk: feature map l2-rel {err:.3e}")
    # 5e-2, not 2e-2: scripts/exp/resnet_train_bf16_noise.py (pure PyTorch, CPU) measures 4.45e-2 for ANY bf16-operand
    # evaluation of this loop — 53 per-frame re-normalisations amplify the operand rounding
    assert err < 5e-2, err
    ref_bns = [b for b in ref_seq.modules() if isinstance(b, torch.nn.BatchNorm2d)]
    got_bns = [b for b in m.resnet.modules() if isinstance(b, torch.nn.BatchNorm2d)]
    assert len(ref_bns) == len(got_bns) == 53
    worst_m = worst_v = 0.0
    for rb, gb in zip(ref_bns, got_bns):
        assert int(gb.num_batches_tracked) == int(rb.num_batches_tracked) == 6
        worst_m = max(worst_m, ((gb.running_mean.cpu() - rb.running_mean).norm() / rb.running_mean.norm()).item())
        worst_v = max(worst_v, ((gb.running_var.cpu() - rb.running_var).norm() / rb.running_var.norm()).item())
    print(f"train-mode trunk: running_mean worst l2-rel {worst_m:.3e}, running_var {worst_v:.3e}")
    assert worst_m < 5e-2 and worst_v < 5e-2, (worst_m, worst_v)
    # the per-frame statistics make the result independent of how frames are batched: one frame alone (the
    # reference's own call shape) gives the same tile as the same frame inside the batch of six
    m_one = M.ResnetFeatureExtractor().to(dev)
    m_one.load_state_dict(m.state_dict())
    with torch.no_grad():
        tile = m_one.encode(x[1, 2].to(dev))
        again = m(x.to(dev))
    d = ((tile - again[1, :, 0:16, 32:48]).norm() / tile.norm()).item()
    print(f"train-mode trunk: one frame alone vs inside the batch l2-rel {d:.3e}")
    assert d < 1e-5, d


def test_resnet_train_mode_trunk_vs_reference_golden(golden_dir):
    """tests/golden/resnet_train.npz was written by the UNMODIFIED reference class built with its default arguments
    (train-mode trunk, one frame per call): the drop-in, which batches the three frames, must land on the same feature
    map, running statistics and num_batches_tracked."""
    import numpy as np
    import resnet_extractor as M
    dev = _dev()
    warnings.filterwarnings("ignore")
    G = np.load(os.path.join(golden_dir, "resnet_train.npz"))
    torch.manual_seed(0)
    m = M.ResnetFeatureExtractor()
    O.resnet_randomise_bn(m.resnet, 29)
    m = m.to(dev)
    x = torch.rand((1, 3, 3, 48, 64), generator=torch.Generator().manual_seed(63)).to(dev)
    with torch.no_grad():
        y = m(x)
    want = torch.from_numpy(G["y"])
    err = ((y.cpu() - want).norm() / want.norm()).item()
    print(f"train-mode trunk vs reference golden: feature map l2-rel {err:.3e}")
    assert err < 5e-2, err
    bns = [b for b in m.resnet.modules() if isinstance(b, torch.nn.BatchNorm2d)]
    assert [int(b.num_batches_tracked) for b in bns] == list(G["nbt"])
    for tag, b in (("first", bns[0]), ("last", bns[-1])):
        for name in ("running_mean", "running_var"):
            ref = torch.from_numpy(G[f"{tag}/{name}"])
            e = ((getattr(b, name).cpu() - ref).norm() / ref.norm()).item()
            print(f"train-mode trunk vs reference golden: {tag} BatchNorm {name} l2-rel {e:.3e}")
            assert e < 5e-2, (tag, name, e)


def test_video_processor_restatement_and_il_step():
    """video_processor.VideoProcessor (ABSENT from the reference: restated from its call sites, parity
    unpinned) against the oracle's restatement, then one imitation-learning forward of
    rovr/imitation_learning.py:72-87 end to end on the drop-ins: frames -> VideoProcessor -> PolicyNetwork2UNet."""
    import copy
    from video_processor import VideoProcessor
    from policy_net_2 import PolicyNetwork2UNet
    dev = _dev()
    warnings.filterwarnings("ignore")
    torch.manual_seed(3)
    vp = VideoProcessor()
    O.resnet_randomise_bn(vp.resnet, 61)
    ref_seq = copy.deepcopy(vp.resnet).eval()
    lw, lb = vp.linear.weight.detach().clone(), vp.linear.bias.detach().clone()
    vp = vp.to(dev)
    g = torch.Generator().manual_seed(8)
    frames = torch.rand((1, 20, 3, 64, 80), generator=g)
    enc, flat = vp(frames.to(dev))
    assert enc.shape == (1, 1, 160, 160) and flat.shape == (1, 20, 1024)
    with torch.no_grad():
        enc_ref, flat_ref = O.video_processor_forward(ref_seq, lw, lb, frames)
    assert _rel(flat, flat_ref) < 2e-2 and _rel(enc, enc_ref) < 2e-2
    assert float(enc[0, 0, 128:, :].abs().sum()) == 0.0                       # tiles 20..24 stay empty
    assert torch.equal(enc[0, 0, 32:64, 64:96].reshape(-1), flat[0, 7])       # frame 7 -> tile (1, 2)
    # rovr/rovr.py:200: overwrite one tile with the encoding of a reconstructed frame
    enc2 = vp.insert_encoded_frame_batch(torch.tensor(3).view(-1, 1), frames[0, 5:6].to(dev), enc.detach().clone())
    assert _rel(enc2[0, 0, 0:32, 96:128].reshape(-1), flat[0, 5]) < 1e-3
    # gradients reach the projection only
    (flat ** 2).sum().backward()
    assert vp.linear.weight.grad is not None and all(p.grad is None for p in vp.resnet.parameters())
    # imitation_learning.py:82-87
    pn2 = PolicyNetwork2UNet().to(dev).train()
    encoded = torch.stack([enc.detach()] * 20, dim=0).squeeze(1)             # [20, 1, 160, 160]
    out = pn2(encoded, flat.detach().transpose(0, 1), torch.arange(20, device=dev).view(20, 1, 1), extra=True)
    assert out.shape == (20, 20) and torch.isfinite(out).all()
