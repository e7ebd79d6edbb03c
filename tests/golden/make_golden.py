"""Generate golden fixtures from the UNMODIFIED reference modules.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It imports the reference classes from /root/reference/rovr, loads the deterministic weights of
oracle.rovr_oracle into them with load_state_dict(strict=True) (which also proves the oracle's
layer specs match the reference's state_dict keys and shapes), runs them on seeded inputs on the
CPU in fp32 and stores outputs and gradients as small .npz files next to this script. The GPU box
has no /root/reference: tests only read the committed .npz files.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference/rovr")

import rovr_oracle as O  # noqa: E402

torch.set_num_threads(8)
torch.backends.mkldnn.enabled = True


def quiet(fn, *a, **k):
    """The reference forward() bodies print shapes; keep the console clean."""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def sample_idx(numel, k, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randperm(numel, generator=g)[: min(k, numel)]


def grad_digest(name, g, out, k=2048):
    """Store small gradients whole, large ones as (norm, sampled entries)."""
    flat = g.detach().reshape(-1).double()
    out[f"gnorm/{name}"] = np.array(float(flat.norm()))
    if flat.numel() <= 4096:
        out[f"gfull/{name}"] = g.detach().numpy().astype(np.float32)
    else:
        idx = sample_idx(flat.numel(), k, 99)
        out[f"gidx/{name}"] = idx.numpy().astype(np.int64)
        out[f"gval/{name}"] = flat[idx].numpy().astype(np.float32)


def gen_localnet():
    from local_net import LocalNetworkUNetNorm
    sd = O.localnet_state_dict(0)
    net = LocalNetworkUNetNorm()
    assert list(net.state_dict().keys()) == list(sd.keys()), "LocalNet state_dict key order differs"
    net.load_state_dict(sd, strict=True)
    out = {}
    for tag, (B, H, W) in {"a": (2, 32, 32), "b": (1, 64, 40)}.items():
        x, ctx, tgt = O.synthetic_localnet_batch(B, H, W, seed=1234)
        net.zero_grad()
        y = quiet(net, x, ctx)
        loss = torch.nn.MSELoss()(y, tgt)
        loss.backward()
        out[f"{tag}/y"] = y.detach().numpy()
        out[f"{tag}/loss"] = np.array(float(loss.detach()))
        for n, p in net.named_parameters():
            if p.grad is None:
                out[f"{tag}/nograd/{n}"] = np.array(1)
            else:
                grad_digest(f"{tag}/{n}", p.grad, out)
    out["param_order"] = np.array([n for n, _ in net.named_parameters()])
    out["state_keys"] = np.array(list(net.state_dict().keys()))
    np.savez_compressed(os.path.join(HERE, "localnet.npz"), **out)
    print("localnet.npz", len(out), "arrays")


def pn1_inputs(b, seed):
    g = torch.Generator().manual_seed(seed)
    image = torch.rand((b, 3, 80, 80), generator=g)
    context = torch.rand((b, 3, 80, 80), generator=g)
    action = torch.randint(0, 25, (b,), generator=g)
    return image, context, action


def gen_pn1():
    from policy_net_1 import PolicyNetwork1UNet
    out = {}
    # actor: logprob at b=5 (fwd+bwd), forward at b=1
    sd = O.pn1_state_dict(0, is_critic=False)
    net = PolicyNetwork1UNet(is_critic=False)
    assert list(net.state_dict().keys()) == list(sd.keys()), "PN1 state_dict key order differs"
    net.load_state_dict(sd, strict=True)
    net.train()
    image, context, action = pn1_inputs(5, 11)
    net.zero_grad()
    torch.manual_seed(777)
    lp = quiet(net.logprob, image, context, action)
    lp.sum().backward()
    out["actor/logprob"] = lp.detach().numpy()
    for n, p in net.named_parameters():
        grad_digest(f"actor/{n}", p.grad, out)
    for n, bfr in net.named_buffers():
        out[f"actor/buf/{n}"] = bfr.detach().clone().numpy()
    net.load_state_dict(sd, strict=True)
    image1, context1, _ = pn1_inputs(1, 12)
    # b = 1 in train mode: BatchNorm over H*W only; conv9/bn9 output has > 1 value per channel
    torch.manual_seed(778)
    idx, logp = quiet(net, image1, context1)
    out["actor/fwd_idx"] = idx.numpy()
    out["actor/fwd_logp"] = logp.numpy()
    # critic
    sdc = O.pn1_state_dict(0, is_critic=True)
    crit = PolicyNetwork1UNet(is_critic=True)
    crit.load_state_dict(sdc, strict=True)
    crit.train()
    crit.zero_grad()
    v = quiet(crit, image, context)
    (v ** 2).sum().backward()
    out["critic/value"] = v.detach().numpy()
    for n, p in crit.named_parameters():
        grad_digest(f"critic/{n}", p.grad, out)
    out["param_order"] = np.array([n for n, _ in net.named_parameters()])
    np.savez_compressed(os.path.join(HERE, "pn1.npz"), **out)
    print("pn1.npz", len(out), "arrays")


def pn2_inputs(b, seed):
    g = torch.Generator().manual_seed(seed)
    enc = torch.rand((b, 1, 160, 160), generator=g)
    feat = torch.randn((b, 1, 1024), generator=g)
    target = torch.randint(0, 20, (b, 1, 1), generator=g)
    a0 = torch.randint(0, 20, (b,), generator=g)
    a1 = (a0 + 1 + torch.randint(0, 19, (b,), generator=g)) % 20
    return enc, feat, target, torch.stack([a0, a1], 1)


def gen_pn2():
    from policy_net_2 import PolicyNetwork2UNet
    out = {}
    sd = O.pn2_state_dict(0, is_critic=False)
    net = PolicyNetwork2UNet(is_critic=False)
    assert list(net.state_dict().keys()) == list(sd.keys()), "PN2 state_dict key order differs"
    net.load_state_dict(sd, strict=True)
    net.train()
    enc, feat, target, action = pn2_inputs(20, 21)
    # imitation-learning entry (rovr/imitation_learning.py:87): extra=True -> masked logits
    net.zero_grad()
    logits = quiet(net, enc, feat, target, extra=True)
    (logits ** 2).sum().backward()
    out["actor/il_logits"] = logits.detach().numpy()
    for n, p in net.named_parameters():
        if p.grad is None:
            out[f"actor/il/nograd/{n}"] = np.array(1)
        else:
            grad_digest(f"actor/il/{n}", p.grad, out)
    for n, bfr in net.named_buffers():
        out[f"actor/il/buf/{n}"] = bfr.detach().clone().numpy()
    # PPO logprob (rovr/rovr.py:312): image [T,160,160], context [T,1024], target [T,1]
    net.load_state_dict(sd, strict=True)
    net.zero_grad()
    torch.manual_seed(779)
    lp = quiet(net.logprob, enc[:, 0], feat[:, 0], target[:, 0], action, torch.device("cpu"))
    lp.sum().backward()
    out["actor/logprob"] = lp.detach().numpy()
    for n, p in net.named_parameters():
        if p.grad is not None:
            grad_digest(f"actor/lp/{n}", p.grad, out)
    # rollout actor forward at b = 1 (rovr/rovr.py:141)
    net.load_state_dict(sd, strict=True)
    torch.manual_seed(780)
    idx, logp = quiet(net, enc[:1], feat[:1], target[:1])
    out["actor/fwd_idx"] = idx.numpy()
    out["actor/fwd_logp"] = logp.numpy()
    # critic at b = T (rovr/rovr.py:299,311)
    sdc = O.pn2_state_dict(0, is_critic=True)
    crit = PolicyNetwork2UNet(is_critic=True)
    crit.load_state_dict(sdc, strict=True)
    crit.train()
    crit.zero_grad()
    v = quiet(crit, enc[:, 0], feat[:, 0], target[:, 0])
    (v ** 2).sum().backward()
    out["critic/value"] = v.detach().numpy()
    for n, p in crit.named_parameters():
        if p.grad is not None:
            grad_digest(f"critic/{n}", p.grad, out)
    out["param_order"] = np.array([n for n, _ in net.named_parameters()])
    np.savez_compressed(os.path.join(HERE, "pn2.npz"), **out)
    print("pn2.npz", len(out), "arrays")


def block_state_dict(module, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in module.state_dict().items():
        if "layer_norm" in k and k.endswith("weight"):
            sd[k] = 1.0 + (torch.rand(v.shape, generator=g) - 0.5) * 0.2
        elif v.dim() >= 2:
            sd[k] = (torch.rand(v.shape, generator=g) - 0.5) * 2.0 * (3.0 / v.shape[-1]) ** 0.5
        else:
            sd[k] = (torch.rand(v.shape, generator=g) - 0.5) * 0.2
    return sd


def gen_common():
    import common_layers as CL
    out = {}
    E, heads, S, T, B = 128, 4, 16, 8, 2
    g = torch.Generator().manual_seed(31)
    x = torch.randn((B, S, E), generator=g)
    enc = torch.randn((B, T, E), generator=g)
    for name, ctor, args in [
        ("self_attn", lambda: CL.SelfAttentionBlock(E, heads, 0.0), (x,)),
        ("cross_attn", lambda: CL.CrossAttentionBlock(E, heads, 0.0), (x, enc)),
        ("ffn", lambda: CL.FeedForwardBlock(E, 0.0), (x,)),
        ("encoder", lambda: CL.EncoderBlock(E, heads, 0.0), (x,)),
        ("decoder", lambda: CL.DecoderBlock(E, heads, 0.0), (x, enc)),
    ]:
        m = ctor()
        sd = block_state_dict(m, 41)
        m.load_state_dict(sd, strict=True)
        m.eval()
        ins = [t.clone().requires_grad_(True) for t in args]
        y = m(*ins)
        (y ** 2).sum().backward()
        out[f"{name}/y"] = y.detach().numpy()
        for i, t in enumerate(ins):
            out[f"{name}/gin{i}"] = t.grad.numpy()
        for n, p in m.named_parameters():
            grad_digest(f"{name}/{n}", p.grad, out)
        out[f"{name}/keys"] = np.array(list(sd.keys()))
    # positional encodings: tokens = P^2 * C with P = 4, C = 8 -> 128
    ipe = CL.ImagePositionalEncoding(4, 4, 8)
    sd = block_state_dict(ipe, 42)
    ipe.load_state_dict(sd, strict=True)
    xi = torch.randn((B, 16, 128), generator=g)
    out["ipe/y"] = ipe(xi).detach().numpy()
    cpe = CL.ContextPositionalEncoding(2, 4, 8, 2)
    sd = block_state_dict(cpe, 43)
    cpe.load_state_dict(sd, strict=True)
    xc = torch.randn((B, 8, 128), generator=g)
    out["cpe/y"] = cpe(xc).detach().numpy()
    np.savez_compressed(os.path.join(HERE, "common_layers.npz"), **out)
    print("common_layers.npz", len(out), "arrays")


def gen_action_lstm():
    from action_lstm import ActionLSTM
    out = {}
    m = ActionLSTM(64, 1, 2)
    sd = block_state_dict(m, 51)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(52)
    for step in range(2):
        action = torch.randint(0, 48, (2, 3), generator=g)
        new_tensor = torch.rand((2, 3, 3, 16, 16), generator=g)
        y = m(action, new_tensor)
        out[f"step{step}/y"] = y.detach().numpy()
        out[f"step{step}/hx"] = m.hx.detach().numpy()
    out["keys"] = np.array(list(sd.keys()))
    np.savez_compressed(os.path.join(HERE, "action_lstm.npz"), **out)
    print("action_lstm.npz", len(out), "arrays")


def resnet_reference_module(cls):
    """ResnetFeatureExtractor(pretrained=False) with seeded default init, deterministic BatchNorm
    statistics, trunk in eval mode and frozen (what pretrained=True sets up, rovr/resnet_extractor.py:11-14;
    the pretrained weights themselves need a download)."""
    torch.manual_seed(0)
    m = quiet(cls, pretrained=False)
    O.resnet_randomise_bn(m.resnet, 61)
    m.resnet.eval()
    for p in m.resnet.parameters():
        p.requires_grad = False
    return m


def resnet_inputs():
    g = torch.Generator().manual_seed(62)
    return {"a": torch.rand((1, 3, 3, 48, 64), generator=g),        # Resize up, non-square
            "b": torch.rand((1, 2, 3, 224, 224), generator=g),      # Resize is the identity
            "c": torch.rand((1, 1, 3, 256, 256), generator=g)}      # the dataset's frame size (down-sampling)


def gen_resnet():
    import warnings
    warnings.filterwarnings("ignore")
    from resnet_extractor import ResnetFeatureExtractor
    m = resnet_reference_module(ResnetFeatureExtractor)
    out = {"keys": np.array(list(m.state_dict().keys()))}
    for tag, x in resnet_inputs().items():
        m.zero_grad()
        y = m(x)
        (y ** 2).sum().backward()
        out[f"{tag}/y"] = y.detach().numpy()
        grad_digest(f"{tag}/linear.weight", m.linear.weight.grad, out)
        grad_digest(f"{tag}/linear.bias", m.linear.bias.grad, out)
        # the preprocessing of the first frame (PIL path), to pin the GPU resampler
        out[f"{tag}/prep0"] = m.preprocessing(x[0, 0]).numpy()[:, ::7, ::5].copy()
    np.savez_compressed(os.path.join(HERE, "resnet.npz"), **out)
    print("resnet.npz", len(out), "arrays")


def resnet_train_inputs():
    return torch.rand((1, 3, 3, 48, 64), generator=torch.Generator().manual_seed(63))


def gen_resnet_train():
    """The reference's DEFAULT constructor (`pretrained=False`, rovr/resnet_extractor.py:6-8): nothing puts the trunk
    in eval mode, so every BatchNorm runs on batch statistics — of ONE frame, since `encode` is called per frame
    (:31-33, :42-47) — and its running buffers advance once per frame. Pins that behaviour: feature map, the running
    statistics of the first and the last BatchNorm and `num_batches_tracked` after one forward of a 3-frame clip."""
    import warnings
    warnings.filterwarnings("ignore")
    from resnet_extractor import ResnetFeatureExtractor
    torch.manual_seed(0)
    m = quiet(ResnetFeatureExtractor)                       # default arguments: train-mode, trainable trunk
    O.resnet_randomise_bn(m.resnet, 29)
    assert all(b.training for b in m.resnet.modules() if isinstance(b, torch.nn.BatchNorm2d))
    with torch.no_grad():
        y = m(resnet_train_inputs())
    bns = [b for b in m.resnet.modules() if isinstance(b, torch.nn.BatchNorm2d)]
    out = {"y": y.numpy(), "nbt": np.array([int(b.num_batches_tracked) for b in bns])}
    for tag, b in (("first", bns[0]), ("last", bns[-1])):
        out[f"{tag}/running_mean"] = b.running_mean.numpy().copy()
        out[f"{tag}/running_var"] = b.running_var.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "resnet_train.npz"), **out)
    print("resnet_train.npz", len(out), "arrays")


def gen_lpips():
    """LPIPS(net='vgg'): the `lpips` package is not in this image (parity unpinned against it); what IS
    pinned is the VGG16 trunk — torchvision.models.vgg16().features run as torchvision wrote it, with the
    oracle's seeded weights loaded by their feature indices — plus the published LPIPS head arithmetic
    written out literally here (independent of oracle.lpips_vgg, which the CPU test compares with this)."""
    import warnings
    warnings.filterwarnings("ignore")
    import torchvision
    sd = O.lpips_state_dict(0)
    vgg = torchvision.models.vgg16(weights=None).features.eval()
    tv = {}
    for k, convs in enumerate(O.VGG16_SLICES):
        for idx, _, _ in convs:
            tv[f"{idx}.weight"] = sd[f"net.slice{k + 1}.{idx}.weight"]
            tv[f"{idx}.bias"] = sd[f"net.slice{k + 1}.{idx}.bias"]
    vgg.load_state_dict(tv, strict=True)
    taps = [3, 8, 15, 22, 29]                      # relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 (lpips/pretrained_networks.py)

    def feats(x):
        out = []
        for i, layer in enumerate(vgg):
            x = layer(x)
            if i in taps:
                out.append(x)
            if i == taps[-1]:
                break
        return out

    def lpips(in0, in1, normalize):
        if normalize:
            in0, in1 = 2 * in0 - 1, 2 * in1 - 1
        shift = torch.Tensor([-.030, -.088, -.188])[None, :, None, None]
        scale = torch.Tensor([.458, .448, .450])[None, :, None, None]
        o0, o1 = feats((in0 - shift) / scale), feats((in1 - shift) / scale)
        val = 0
        for k in range(5):
            a = o0[k] / (torch.sqrt(torch.sum(o0[k] ** 2, dim=1, keepdim=True)) + 1e-10)
            b = o1[k] / (torch.sqrt(torch.sum(o1[k] ** 2, dim=1, keepdim=True)) + 1e-10)
            val = val + torch.nn.functional.conv2d((a - b) ** 2, sd[f"lin{k}.model.1.weight"]).mean([2, 3], keepdim=True)
        return val

    out = {"keys": np.array(list(sd.keys()))}
    for tag, (n, h, w, normalize) in {"a": (2, 32, 32, False), "b": (1, 64, 48, True)}.items():
        g = torch.Generator().manual_seed(71 + n)
        in0 = torch.rand((n, 3, h, w), generator=g).requires_grad_(True)
        in1 = torch.rand((n, 3, h, w), generator=g)
        val = lpips(in0, in1, normalize)
        val.mean().backward()
        out[f"{tag}/val"] = val.detach().numpy()
        out[f"{tag}/grad_in0"] = in0.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "lpips.npz"), **out)
    print("lpips.npz", len(out), "arrays")


if __name__ == "__main__":
    which = sys.argv[1:] or ["localnet", "pn1", "pn2", "common", "action_lstm", "resnet", "resnet_train", "lpips"]
    for w in which:
        globals()["gen_" + w]()
