"""CPU: the oracle (oracle/rovr_oracle.py) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py). This is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

import rovr_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _check_grads(G, prefix, grads, rtol=2e-4):
    n_checked = 0
    for name, g in grads.items():
        if g is None:
            continue
        key = f"{prefix}/{name}"
        gn = float(G[f"gnorm/{key}"])
        assert abs(float(g.double().norm()) - gn) <= rtol * max(gn, 1e-6), name
        if f"gfull/{key}" in G:
            ref = torch.from_numpy(G[f"gfull/{key}"])
            assert torch.allclose(g, ref, rtol=rtol, atol=rtol * max(gn, 1e-6) / max(1, g.numel()) ** 0.5), name
        else:
            idx = torch.from_numpy(G[f"gidx/{key}"])
            ref = torch.from_numpy(G[f"gval/{key}"])
            got = g.reshape(-1)[idx]
            assert torch.allclose(got, ref, rtol=rtol, atol=rtol * max(gn, 1e-6) / g.numel() ** 0.5), name
        n_checked += 1
    return n_checked


@pytest.mark.parametrize("tag,shape", [("a", (2, 32, 32)), ("b", (1, 64, 40))])
def test_localnet_matches_reference(golden_dir, tag, shape):
    G = _load(golden_dir, "localnet.npz")
    sd = O.localnet_state_dict(0)
    x, ctx, tgt = O.synthetic_localnet_batch(*shape, seed=1234)
    y, loss, grads = O.localnet_step(sd, x, ctx, tgt)
    assert np.allclose(y.numpy(), G[f"{tag}/y"], rtol=1e-5, atol=1e-6)
    assert abs(float(loss) - float(G[f"{tag}/loss"])) < 1e-6
    assert _check_grads(G, tag, grads) == 22
    # the reference leaves the 10 BatchNorm2d without gradients
    nograd = [k for k in G.files if k.startswith(f"{tag}/nograd/")]
    assert len(nograd) == 20


def test_localnet_state_dict_layout(golden_dir):
    G = _load(golden_dir, "localnet.npz")
    sd = O.localnet_state_dict(0)
    assert list(G["state_keys"]) == list(sd.keys())
    assert len(sd) == 72
    assert sum(v.numel() for k, v in sd.items() if k in O.LOCALNET_LIVE) == 3791939


def _pn1_inputs(b, seed):
    g = torch.Generator().manual_seed(seed)
    image = torch.rand((b, 3, 80, 80), generator=g)
    context = torch.rand((b, 3, 80, 80), generator=g)
    action = torch.randint(0, 25, (b,), generator=g)
    return image, context, action


def test_pn1_matches_reference(golden_dir):
    G = _load(golden_dir, "pn1.npz")
    sd = O.pn1_state_dict(0, False)
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
            for k, v in sd.items()}
    image, context, action = _pn1_inputs(5, 11)
    torch.manual_seed(777)
    expo = torch.empty((5, 25)).exponential_()
    lp = O.pn1_logprob(leaf, image, context, action, expo)
    lp.sum().backward()
    assert np.allclose(lp.detach().numpy(), G["actor/logprob"], rtol=2e-4, atol=1e-5)
    grads = {k: v.grad for k, v in leaf.items() if v.requires_grad}
    assert _check_grads(G, "actor", grads, rtol=2e-3) == len(grads)
    # b = 1 rollout entry: indices bit-exact
    image1, context1, _ = _pn1_inputs(1, 12)
    torch.manual_seed(778)
    expo1 = torch.empty((1, 25)).exponential_()
    idx, logp = O.pn1_forward(sd, image1, context1, False, expo1)
    assert np.array_equal(idx.numpy(), G["actor/fwd_idx"])
    assert np.allclose(logp.numpy(), G["actor/fwd_logp"], rtol=1e-4, atol=1e-5)
    # critic
    sdc = O.pn1_state_dict(0, True)
    leafc = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
             for k, v in sdc.items()}
    v = O.pn1_forward(leafc, image, context, True)
    (v ** 2).sum().backward()
    assert np.allclose(v.detach().numpy(), G["critic/value"], rtol=2e-4, atol=1e-5)
    gradsc = {k: t.grad for k, t in leafc.items() if t.requires_grad}
    assert _check_grads(G, "critic", gradsc, rtol=2e-3) == len(gradsc)


def test_pn1_running_stats(golden_dir):
    """train-mode BatchNorm buffer update (momentum 0.1, unbiased variance) on the first layer."""
    G = _load(golden_dir, "pn1.npz")
    sd = O.pn1_state_dict(0, False)
    image, context, _ = _pn1_inputs(5, 11)
    import torch.nn.functional as F
    pre = F.conv2d(torch.cat([image, context], 1), sd["conv1.weight"], sd["conv1.bias"], padding=1)
    rm, rv = O.bn_running_update(pre, sd["bn1.running_mean"], sd["bn1.running_var"])
    assert np.allclose(rm.numpy(), G["actor/buf/bn1.running_mean"], rtol=1e-4, atol=1e-6)
    assert np.allclose(rv.numpy(), G["actor/buf/bn1.running_var"], rtol=1e-4, atol=1e-6)
    assert int(G["actor/buf/bn1.num_batches_tracked"]) == 1


def _pn2_inputs(b, seed):
    g = torch.Generator().manual_seed(seed)
    enc = torch.rand((b, 1, 160, 160), generator=g)
    feat = torch.randn((b, 1, 1024), generator=g)
    target = torch.randint(0, 20, (b, 1, 1), generator=g)
    a0 = torch.randint(0, 20, (b,), generator=g)
    a1 = (a0 + 1 + torch.randint(0, 19, (b,), generator=g)) % 20
    return enc, feat, target, torch.stack([a0, a1], 1)


def _leaf(sd):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
            for k, v in sd.items()}


def test_pn2_matches_reference(golden_dir):
    G = _load(golden_dir, "pn2.npz")
    sd = O.pn2_state_dict(0, False)
    enc, feat, target, action = _pn2_inputs(20, 21)
    leaf = _leaf(sd)
    logits = O.pn2_forward(leaf, enc, feat, target, False, extra=True)
    (logits ** 2).sum().backward()
    assert np.allclose(logits.detach().numpy(), G["actor/il_logits"], rtol=5e-4, atol=5e-5)
    grads = {k: v.grad for k, v in leaf.items() if v.requires_grad and v.grad is not None}
    assert _check_grads(G, "actor/il", grads, rtol=5e-3) == len(grads)
    # context_conv is constructed but unused: no gradient in the reference either
    assert all(leaf[k].grad is None for k in leaf if k.startswith("context_conv") and leaf[k].requires_grad)
    # PPO logprob
    leaf = _leaf(sd)
    torch.manual_seed(779)
    expo = torch.empty((20, 20)).exponential_()
    lp = O.pn2_logprob(leaf, enc[:, 0], feat[:, 0], target[:, 0], action, expo)
    lp.sum().backward()
    assert np.allclose(lp.detach().numpy(), G["actor/logprob"], rtol=5e-4, atol=5e-5)
    grads = {k: v.grad for k, v in leaf.items() if v.requires_grad and v.grad is not None}
    assert _check_grads(G, "actor/lp", grads, rtol=5e-3) == len(grads)
    # rollout forward b = 1: indices bit-exact
    torch.manual_seed(780)
    expo1 = torch.empty((1, 20)).exponential_()
    idx, logp = O.pn2_forward(sd, enc[:1], feat[:1], target[:1], False, expo=expo1)
    assert np.array_equal(idx.numpy(), G["actor/fwd_idx"])
    assert np.allclose(logp.numpy(), G["actor/fwd_logp"], rtol=1e-4, atol=1e-5)
    # critic
    sdc = O.pn2_state_dict(0, True)
    leafc = _leaf(sdc)
    v = O.pn2_forward(leafc, enc[:, 0], feat[:, 0], target[:, 0], True)
    (v ** 2).sum().backward()
    assert np.allclose(v.detach().numpy(), G["critic/value"], rtol=5e-4, atol=5e-5)
    gradsc = {k: t.grad for k, t in leafc.items() if t.requires_grad and t.grad is not None}
    assert _check_grads(G, "critic", gradsc, rtol=5e-3) == len(gradsc)


def _block_sd(G, name, shapes_from):
    return None


def test_common_layers_match_reference(golden_dir):
    G = _load(golden_dir, "common_layers.npz")
    E, heads, S, T, B = 128, 4, 16, 8, 2
    g = torch.Generator().manual_seed(31)
    x = torch.randn((B, S, E), generator=g)
    enc = torch.randn((B, T, E), generator=g)

    def make_sd(keys, shapes, seed):
        gg = torch.Generator().manual_seed(seed)
        sd = {}
        for k in keys:
            shp = shapes[k]
            if "layer_norm" in k and k.endswith("weight"):
                sd[k] = 1.0 + (torch.rand(shp, generator=gg) - 0.5) * 0.2
            elif len(shp) >= 2:
                sd[k] = (torch.rand(shp, generator=gg) - 0.5) * 2.0 * (3.0 / shp[-1]) ** 0.5
            else:
                sd[k] = (torch.rand(shp, generator=gg) - 0.5) * 0.2
        return sd

    def shapes_for(prefixes):
        s = {}
        for p, kind in prefixes:
            if kind == "attn":
                s[p + "attention.in_proj_weight"] = (3 * E, E)
                s[p + "attention.in_proj_bias"] = (3 * E,)
                s[p + "attention.out_proj.weight"] = (E, E)
                s[p + "attention.out_proj.bias"] = (E,)
                s[p + "layer_norm.weight"] = (E,)
                s[p + "layer_norm.bias"] = (E,)
            elif kind == "xattn":
                s[p + "attention.in_proj_weight"] = (3 * E, E)
                s[p + "attention.in_proj_bias"] = (3 * E,)
                s[p + "attention.out_proj.weight"] = (E, E)
                s[p + "attention.out_proj.bias"] = (E,)
                s[p + "layer_norm.weight"] = (E,)
                s[p + "layer_norm.bias"] = (E,)
                s[p + "layer_norm_encoder_output.weight"] = (E,)
                s[p + "layer_norm_encoder_output.bias"] = (E,)
            else:
                s[p + "fc1.weight"] = (E // 4, E)
                s[p + "fc1.bias"] = (E // 4,)
                s[p + "fc2.weight"] = (E, E // 4)
                s[p + "fc2.bias"] = (E,)
                s[p + "layer_norm.weight"] = (E,)
                s[p + "layer_norm.bias"] = (E,)
        return s

    cases = {
        "self_attn": ([("", "attn")], lambda sd, a: O.self_attention_block(sd, "", a[0], heads), (x,)),
        "cross_attn": ([("", "xattn")], lambda sd, a: O.cross_attention_block(sd, "", a[0], a[1], heads), (x, enc)),
        "ffn": ([("", "ffn")], lambda sd, a: O.feed_forward_block(sd, "", a[0]), (x,)),
        "encoder": ([("attention.", "attn"), ("feed_forward.", "ffn")],
                    lambda sd, a: O.encoder_block(sd, "", a[0], heads), (x,)),
        "decoder": ([("attention.", "attn"), ("cross_attention.", "xattn"), ("feed_forward.", "ffn")],
                    lambda sd, a: O.decoder_block(sd, "", a[0], a[1], heads), (x, enc)),
    }
    for name, (prefixes, fn, args) in cases.items():
        keys = list(G[f"{name}/keys"])
        shapes = shapes_for(prefixes)
        assert sorted(keys) == sorted(shapes.keys()), name
        sd = {k: v.requires_grad_(True) for k, v in make_sd(keys, shapes, 41).items()}
        ins = [t.clone().requires_grad_(True) for t in args]
        y = fn(sd, ins)
        (y ** 2).sum().backward()
        assert np.allclose(y.detach().numpy(), G[f"{name}/y"], rtol=2e-4, atol=2e-5), name
        for i, t in enumerate(ins):
            assert np.allclose(t.grad.numpy(), G[f"{name}/gin{i}"], rtol=2e-3, atol=2e-4), (name, i)
        grads = {k: v.grad for k, v in sd.items()}
        assert _check_grads(G, name, grads, rtol=2e-3) == len(grads)


def test_positional_encodings_match_reference(golden_dir):
    G = _load(golden_dir, "common_layers.npz")
    g = torch.Generator().manual_seed(31)
    _ = torch.randn((2, 16, 128), generator=g)
    _ = torch.randn((2, 8, 128), generator=g)
    xi = torch.randn((2, 16, 128), generator=g)

    def lin_sd(names, seed):
        gg = torch.Generator().manual_seed(seed)
        sd = {}
        for n in names:
            sd[n + ".weight"] = (torch.rand((128, 1), generator=gg) - 0.5) * 2.0 * 3.0 ** 0.5
            sd[n + ".bias"] = (torch.rand((128,), generator=gg) - 0.5) * 0.2
        return sd

    y = O.image_positional_encoding(lin_sd(["positional_encoder"], 42), "", xi, 4)
    assert np.allclose(y.numpy(), G["ipe/y"], rtol=1e-5, atol=1e-5)
    xc = torch.randn((2, 8, 128), generator=g)
    y = O.context_positional_encoding(
        lin_sd(["patch_positional_encoder", "context_positional_encoder"], 43), "", xc, 2, 2)
    assert np.allclose(y.numpy(), G["cpe/y"], rtol=1e-5, atol=1e-5)


def test_action_lstm_matches_reference(golden_dir):
    G = _load(golden_dir, "action_lstm.npz")
    keys = list(G["keys"])
    shapes = {"lstm.weight_ih": (256, 2307), "lstm.weight_hh": (256, 64), "lstm.bias_ih": (256,),
              "lstm.bias_hh": (256,), "fc.weight": (19200, 64), "fc.bias": (19200,)}
    assert keys == list(shapes.keys())
    gg = torch.Generator().manual_seed(51)
    sd = {}
    for k in keys:
        shp = shapes[k]
        if len(shp) >= 2:
            sd[k] = (torch.rand(shp, generator=gg) - 0.5) * 2.0 * (3.0 / shp[-1]) ** 0.5
        else:
            sd[k] = (torch.rand(shp, generator=gg) - 0.5) * 0.2
    g = torch.Generator().manual_seed(52)
    hx = torch.zeros(2, 64)
    cx = torch.zeros(2, 64)
    for step in range(2):
        action = torch.randint(0, 48, (2, 3), generator=g)
        new_tensor = torch.rand((2, 3, 3, 16, 16), generator=g)
        y, hx, cx = O.action_lstm_step(sd, action, new_tensor, hx, cx)
        assert np.allclose(y.numpy(), G[f"step{step}/y"], rtol=1e-4, atol=1e-5)
        assert np.allclose(hx.numpy(), G[f"step{step}/hx"], rtol=1e-4, atol=1e-5)


def _resnet_oracle_module():
    """Same recipe as tests/golden/make_golden.py::resnet_reference_module, built from torchvision
    directly (the reference file only wraps it): seeded init, deterministic BN statistics, eval."""
    import warnings
    import torchvision.models as models
    warnings.filterwarnings("ignore")
    torch.manual_seed(0)
    net = models.resnet50(pretrained=False)
    linear = torch.nn.Linear(2048, 768)
    seq = torch.nn.Sequential(*(list(net.children())[:-1]))
    O.resnet_randomise_bn(seq, 61)
    seq.eval()
    return seq, linear


def _resnet_inputs():
    g = torch.Generator().manual_seed(62)
    return {"a": torch.rand((1, 3, 3, 48, 64), generator=g), "b": torch.rand((1, 2, 3, 224, 224), generator=g),
            "c": torch.rand((1, 1, 3, 256, 256), generator=g)}


def test_resnet_extractor_matches_reference(golden_dir):
    G = _load(golden_dir, "resnet.npz")
    seq, linear = _resnet_oracle_module()
    x = _resnet_inputs()["a"]
    with torch.no_grad():
        y = O.resnet_extractor_forward(seq, linear.weight, linear.bias, x)
    assert np.allclose(y.numpy(), G["a/y"], rtol=1e-4, atol=1e-5)
    keys = ["resnet." + k for k in seq.state_dict().keys()] + ["linear.weight", "linear.bias"]
    assert keys == list(G["keys"])


def test_resnet_train_mode_oracle_matches_reference(golden_dir):
    """The reference's default constructor leaves the trunk in training mode and encodes one frame per call
    (rovr/resnet_extractor.py:6-8,31-33,42-47): per-frame batch statistics, running buffers advanced once per frame.
    The oracle's loop on a train-mode trunk must reproduce tests/golden/resnet_train.npz (made by the imported
    reference class): feature map, first / last BatchNorm running statistics, num_batches_tracked."""
    import warnings
    import torchvision.models as models
    warnings.filterwarnings("ignore")
    G = _load(golden_dir, "resnet_train.npz")
    torch.manual_seed(0)
    net = models.resnet50(pretrained=False)
    linear = torch.nn.Linear(2048, 768)
    seq = torch.nn.Sequential(*(list(net.children())[:-1]))
    O.resnet_randomise_bn(seq, 29)
    seq.train()
    x = torch.rand((1, 3, 3, 48, 64), generator=torch.Generator().manual_seed(63))
    with torch.no_grad():
        y = O.resnet_extractor_forward(seq, linear.weight, linear.bias, x)
    assert np.allclose(y.numpy(), G["y"], rtol=1e-4, atol=1e-5)
    bns = [b for b in seq.modules() if isinstance(b, torch.nn.BatchNorm2d)]
    assert [int(b.num_batches_tracked) for b in bns] == list(G["nbt"]) == [3] * 53
    for tag, b in (("first", bns[0]), ("last", bns[-1])):
        assert np.allclose(b.running_mean.numpy(), G[f"{tag}/running_mean"], rtol=1e-4, atol=1e-6)
        assert np.allclose(b.running_var.numpy(), G[f"{tag}/running_var"], rtol=1e-4, atol=1e-6)
    # and the statistics really are per frame: the same frames in ONE batch of three give another result
    seq2 = torch.nn.Sequential(*(list(models.resnet50(pretrained=False).children())[:-1]))
    seq2.load_state_dict({k: v for k, v in seq.state_dict().items()})
    import torchvision.transforms as T
    prep = T.Compose([T.ToPILImage(), T.Resize((224, 224)), T.ToTensor()])
    with torch.no_grad():
        batched = seq2.train()(torch.stack([prep(f) for f in x[0]]))
        tiles = torch.nn.functional.linear(batched.flatten(1), linear.weight, linear.bias).view(3, 3, 16, 16)
    assert not np.allclose(tiles[1].numpy(), G["y"][0, :, 0:16, 16:32], rtol=1e-2, atol=1e-3)


def test_lpips_oracle_matches_torchvision_golden(golden_dir):
    """oracle.lpips_vgg vs tests/golden/lpips.npz (torchvision's own VGG16 `features` + the published LPIPS
    head written out literally by make_golden.py). The `lpips` package itself is absent: parity unpinned
    against it, pinned against torchvision's VGG16 structure."""
    G = np.load(os.path.join(golden_dir, "lpips.npz"))
    sd = O.lpips_state_dict(0)
    assert list(sd.keys()) == list(G["keys"])
    for tag, (n, h, w, nz) in {"a": (2, 32, 32, False), "b": (1, 64, 48, True)}.items():
        g = torch.Generator().manual_seed(71 + n)
        in0 = torch.rand((n, 3, h, w), generator=g).requires_grad_(True)
        in1 = torch.rand((n, 3, h, w), generator=g)
        v = O.lpips_vgg(sd, in0, in1, nz)
        v.mean().backward()
        assert v.shape == (n, 1, 1, 1)
        assert np.allclose(v.detach().numpy(), G[f"{tag}/val"], rtol=1e-5, atol=1e-8)
        assert np.allclose(in0.grad.numpy(), G[f"{tag}/grad_in0"], rtol=1e-4, atol=1e-8)
    # identical images have distance 0; the distance is symmetric
    x = torch.rand((1, 3, 32, 32), generator=torch.Generator().manual_seed(3))
    y = torch.rand((1, 3, 32, 32), generator=torch.Generator().manual_seed(4))
    assert float(O.lpips_vgg(sd, x, x)) == 0.0
    assert abs(float(O.lpips_vgg(sd, x, y)) - float(O.lpips_vgg(sd, y, x))) < 1e-7
