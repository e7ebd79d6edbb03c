"""CPU oracle for the ROVR hot path — TEST INFRASTRUCTURE ONLY.

A plain PyTorch fp32 *functional* restatement of the reference networks' arithmetic, written
from the reference's forward() bodies (each function cites the lines it follows). It is the
checker the CUDA path is compared against; nothing under oracle/ is ever imported by the product
package. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import it.

Pinning: the reference ships no tests, golden vectors or fixtures, so the
oracle is pinned against the reference itself: tests/golden/make_golden.py imports the unmodified
reference classes from /root/reference/rovr in the build container, loads the deterministic
weights of `make_state_dict` into them, and stores their outputs / gradients under
tests/golden/*.npz. tests/test_oracle_golden.py checks this file against those fixtures.
Third-party arithmetic: torch.nn.functional (torch 2.11.0 in this image; the reference pins
nothing) and torchvision.models.resnet50 (0.26.0) for the frame extractor.

The functions take a `state_dict`-style mapping of tensors so that the same weights can be fed
to the reference classes, to this oracle and to the B200 modules.
"""
import math

import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------
# deterministic weights (shared by the golden generator, the oracle tests and the GPU tests)
# ------------------------------------------------------------------------------------------------
# LocalNetworkUNetNorm: registration order of rovr/local_net.py:12-39
LOCALNET_SPEC = [
    ("conv1", "conv3", 9, 64), ("bn1", "bn", 64), ("conv2", "conv3", 64, 128), ("bn2", "bn", 128),
    ("conv3", "conv3", 128, 256), ("bn3", "bn", 256), ("conv4", "conv3", 256, 512), ("bn4", "bn", 512),
    ("upconv1", "up2", 512, 256), ("bn_up1", "bn", 256), ("conv5", "conv3", 512, 256), ("bn5", "bn", 256),
    ("upconv2", "up2", 256, 128), ("bn_up2", "bn", 128), ("conv6", "conv3", 256, 128), ("bn6", "bn", 128),
    ("upconv3", "up2", 128, 64), ("bn_up3", "bn", 64), ("conv7", "conv3", 128, 64), ("bn7", "bn", 64),
    ("conv8", "conv1", 64, 3),
]
# PolicyNetwork1UNet: rovr/policy_net_1.py:19-57 (fc_final appended per is_critic)
PN1_SPEC = [
    ("conv1", "conv3", 6, 32), ("bn1", "bn", 32), ("conv2", "conv3", 32, 64), ("bn2", "bn", 64),
    ("conv3", "conv3", 64, 128), ("bn3", "bn", 128), ("conv4", "conv3", 128, 256), ("bn4", "bn", 256),
    ("upconv1", "up2", 256, 128), ("bn_up1", "bn", 128), ("conv5", "conv3", 256, 128), ("bn5", "bn", 128),
    ("upconv2", "up2", 128, 64), ("bn_up2", "bn", 64), ("conv6", "conv3", 128, 64), ("bn6", "bn", 64),
    ("upconv3", "up2", 64, 32), ("bn_up3", "bn", 32), ("conv7", "conv3", 64, 32), ("bn7", "bn", 32),
    ("conv8", "conv1", 32, 3), ("bn8", "bn", 3), ("conv9", "conv1", 3, 1), ("bn9", "bn", 1),
]
# PolicyNetwork2UNet: rovr/policy_net_2.py:27-69 (Sequential indices)
PN2_SPEC = [
    ("context_conv.0", "conv3", 3, 128), ("context_conv.3", "conv3", 128, 256),
    ("context_conv.6", "conv3", 256, 512),
    ("video_conv.0", "conv3", 1, 64), ("video_conv.1", "bn", 64),
    ("video_conv.4", "conv3", 64, 128), ("video_conv.5", "bn", 128),
    ("video_conv.8", "conv3", 128, 256), ("video_conv.9", "bn", 256),
    ("video_conv.12", "conv3", 256, 512), ("video_conv.13", "bn", 512),
    ("final_fc.0", "linear", 2048, 1024), ("final_fc.1", "linear", 1024, 512),
    ("final_fc.2", "linear", 512, 256), ("final_fc.3", "linear", 256, 64),
]


def _fill(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def state_dict_from_spec(spec, seed, gain=1.0, bn_affine_noise=True):
    """Deterministic fp32 weights for a layer spec. Conv/linear ~ U(-k, k), k = gain*sqrt(3/fan_in)
    (variance-preserving, so activations stay O(1) through the U-Net); BatchNorm affine slightly
    perturbed around (1, 0) so that its gradients are exercised."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for item in spec:
        name, kind = item[0], item[1]
        if kind == "conv3":
            cin, cout = item[2], item[3]
            k = gain * math.sqrt(3.0 / (cin * 9))
            sd[name + ".weight"] = _fill(gen, (cout, cin, 3, 3), k)
            sd[name + ".bias"] = _fill(gen, (cout,), 0.1)
        elif kind == "conv1":
            cin, cout = item[2], item[3]
            k = gain * math.sqrt(3.0 / cin)
            sd[name + ".weight"] = _fill(gen, (cout, cin, 1, 1), k)
            sd[name + ".bias"] = _fill(gen, (cout,), 0.1)
        elif kind == "up2":
            cin, cout = item[2], item[3]
            k = gain * math.sqrt(3.0 / cin)
            sd[name + ".weight"] = _fill(gen, (cin, cout, 2, 2), k)
            sd[name + ".bias"] = _fill(gen, (cout,), 0.1)
        elif kind == "linear":
            cin, cout = item[2], item[3]
            k = gain * math.sqrt(3.0 / cin)
            sd[name + ".weight"] = _fill(gen, (cout, cin), k)
            sd[name + ".bias"] = _fill(gen, (cout,), 0.1)
        elif kind == "bn":
            c = item[2]
            if bn_affine_noise:
                sd[name + ".weight"] = 1.0 + _fill(gen, (c,), 0.2)
                sd[name + ".bias"] = _fill(gen, (c,), 0.2)
            else:
                sd[name + ".weight"] = torch.ones(c)
                sd[name + ".bias"] = torch.zeros(c)
            sd[name + ".running_mean"] = torch.zeros(c)
            sd[name + ".running_var"] = torch.ones(c)
            sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
        else:
            raise ValueError(kind)
    return sd


def localnet_state_dict(seed=0):
    return state_dict_from_spec(LOCALNET_SPEC, seed, gain=math.sqrt(2.0), bn_affine_noise=False)


def pn1_state_dict(seed=0, is_critic=False):
    sd = state_dict_from_spec(PN1_SPEC, seed, gain=math.sqrt(2.0))
    extra = state_dict_from_spec([("fc_final", "linear", 400, 1 if is_critic else 25)], seed + 1000)
    sd.update(extra)
    return sd


def pn2_state_dict(seed=0, is_critic=False):
    sd = state_dict_from_spec(PN2_SPEC, seed, gain=math.sqrt(2.0))
    extra = state_dict_from_spec([("final_fc.4", "linear", 64, 1 if is_critic else 20)], seed + 1000)
    sd.update(extra)
    return sd


# ------------------------------------------------------------------------------------------------
# synthetic masked-video inputs (mask geometry mirrors rovr/video_ds.py:19,62-87)
# ------------------------------------------------------------------------------------------------
def synthetic_localnet_batch(B, H, W, seed=1234):
    """frame [B,3,H,W], context [B,2,3,H,W], target [B,3,H,W] in [0,1], fp32, CPU.

    A clip is a smooth random field translated a little per frame so that neighbouring frames
    correlate; frame f and its two predecessors are corrupted by a zeroed box (150x100 at 256^2,
    scaled with resolution) whose position follows the raster scan of rovr/video_ds.py:62-87;
    target = clean frame f-1 (rovr/train_local_net_unet.py:44-52)."""
    gen = torch.Generator().manual_seed(seed)
    base = torch.rand((B, 3, 16, 16), generator=gen)
    big = F.interpolate(base, size=(H + 8, W + 8), mode="bilinear", align_corners=False)
    fidx = torch.randint(2, 25, (B,), generator=gen)
    bw, bh = max(1, (150 * W) // 256), max(1, (100 * H) // 256)

    def frame_at(b, n):
        dx, dy = int(n) % 8, (int(n) // 3) % 8
        clean = big[b, :, dy:dy + H, dx:dx + W]
        mask = torch.ones((1, H, W))
        x0 = ((int(n) % 8) * 32 * W) // 256
        y0 = ((int(n) // 8) * (256 // 3) * H) // 256
        mask[:, y0:min(H, y0 + bh), x0:min(W, x0 + bw)] = 0.0
        return clean, clean * mask

    frames, ctx, tgt = [], [], []
    for b in range(B):
        f = int(fidx[b])
        _, cf = frame_at(b, f)
        clean_m, cm = frame_at(b, f - 1)
        _, cn = frame_at(b, f - 2)
        frames.append(cf)
        ctx.append(torch.stack([cn, cm], 0))
        tgt.append(clean_m)
    return (torch.stack(frames).contiguous(), torch.stack(ctx).contiguous(),
            torch.stack(tgt).contiguous())


# ------------------------------------------------------------------------------------------------
# LocalNet
# ------------------------------------------------------------------------------------------------
def localnet_forward(sd, x, context):
    """rovr/local_net.py:46-72. BatchNorm layers are registered (:13-37) but never applied."""
    b = x.shape[0]
    # :48-49  cat along a new frame axis, then fold (frame, channel) -> channel: [x | ctx0 | ctx1]
    h = torch.cat([x[:, None], context], dim=1).reshape(b, 9, x.shape[2], x.shape[3])

    def conv(name, t):
        return F.relu(F.conv2d(t, sd[name + ".weight"], sd[name + ".bias"], padding=1))

    def up(name, t):
        return F.relu(F.conv_transpose2d(t, sd[name + ".weight"], sd[name + ".bias"], stride=2))

    e1 = conv("conv1", h)                                   # :52
    e2 = conv("conv2", F.max_pool2d(e1, 2))                 # :53
    e3 = conv("conv3", F.max_pool2d(e2, 2))                 # :54
    e4 = conv("conv4", F.max_pool2d(e3, 2))                 # :55
    d = conv("conv5", torch.cat([up("upconv1", e4), e3], 1))  # :58-60  (order: [up, skip])
    d = conv("conv6", torch.cat([up("upconv2", d), e2], 1))   # :62-64
    d = conv("conv7", torch.cat([up("upconv3", d), e1], 1))   # :66-68
    return torch.sigmoid(F.conv2d(d, sd["conv8.weight"], sd["conv8.bias"]))  # :71


LOCALNET_LIVE = [n + s for n in ["conv1", "conv2", "conv3", "conv4", "upconv1", "conv5", "upconv2",
                                 "conv6", "upconv3", "conv7", "conv8"] for s in (".weight", ".bias")]


def localnet_step(sd, x, context, target):
    """One training step's arithmetic, rovr/train_local_net_unet.py:102-107,115 restricted to the
    named path: y = net(x, ctx); loss = MSELoss()(y, target); loss.backward() (no LPIPS — it is a
    third-party package that is not part of the path — and no optimizer). Returns y, loss, grads."""
    leaf = {k: (v.clone().requires_grad_(True) if k in LOCALNET_LIVE else v) for k, v in sd.items()}
    y = localnet_forward(leaf, x, context)
    loss = F.mse_loss(y, target)
    loss.backward()
    return y.detach(), loss.detach(), {k: leaf[k].grad for k in LOCALNET_LIVE}


# ------------------------------------------------------------------------------------------------
# Policy network 1
# ------------------------------------------------------------------------------------------------
_BN_EVAL = False   # set by `bn_eval_mode()`: evaluate BatchNorm with the running statistics (module.eval())


class bn_eval_mode:
    """Context manager: the policy-network oracles use eval-mode BatchNorm inside it."""

    def __enter__(self):
        global _BN_EVAL
        self.prev, _BN_EVAL = _BN_EVAL, True

    def __exit__(self, *exc):
        global _BN_EVAL
        _BN_EVAL = self.prev


def _bn_train(sd, name, t, eps=1e-5):
    """Train-mode BatchNorm2d (batch statistics, biased variance for normalisation). Running-stat
    updates are returned by pn*_running_stats, not applied here. Inside `bn_eval_mode()`: eval mode."""
    if _BN_EVAL:
        return F.batch_norm(t, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                            sd[name + ".bias"], False, 0.0, eps)
    return F.batch_norm(t, None, None, sd[name + ".weight"], sd[name + ".bias"], True, 0.0, eps)


def _storage(bf16):
    """(activation round-trip, weight round-trip) emulating where the B200 path stores bf16: the packed
    input, conv / upconv weights, every raw convolution output, every post-ReLU activation and
    every activation gradient. Identity functions when bf16 is False."""
    if not bf16:
        return (lambda t, g=True: t), (lambda p: p)
    return (lambda t, g=True: _RoundTrip.apply(t, g)), \
           (lambda p: p + (p.to(torch.bfloat16).to(torch.float32) - p).detach())


def pn1_unet(sd, x, bf16=False):
    """rovr/policy_net_1.py:60-84: LocalNet-shaped U-Net at half width with BN after every conv.
    bf16=True emulates the bf16 storage points of the B200 path (fp32 arithmetic otherwise)."""
    rt, wq = _storage(bf16)
    x = rt(x, False)

    def cbr(conv, bn, t, pad=1):
        raw = rt(F.conv2d(t, wq(sd[conv + ".weight"]), sd[conv + ".bias"], padding=pad))
        return rt(F.relu(_bn_train(sd, bn, raw)))

    def ubr(up, bn, t):
        raw = rt(F.conv_transpose2d(t, wq(sd[up + ".weight"]), sd[up + ".bias"], stride=2))
        return rt(F.relu(_bn_train(sd, bn, raw)))

    e1 = cbr("conv1", "bn1", x)
    e2 = cbr("conv2", "bn2", F.max_pool2d(e1, 2))
    e3 = cbr("conv3", "bn3", F.max_pool2d(e2, 2))
    e4 = cbr("conv4", "bn4", F.max_pool2d(e3, 2))
    d = cbr("conv5", "bn5", torch.cat([ubr("upconv1", "bn_up1", e4), e3], 1))
    d = cbr("conv6", "bn6", torch.cat([ubr("upconv2", "bn_up2", d), e2], 1))
    d = cbr("conv7", "bn7", torch.cat([ubr("upconv3", "bn_up3", d), e1], 1))
    d = cbr("conv8", "bn8", d, pad=0)                                  # :80
    d = cbr("conv9", "bn9", F.max_pool2d(d, 2), pad=0)                 # :81
    return F.max_pool2d(d, 2)                                          # :82


def pn1_compute_logits(sd, image, context, bf16=False):
    """rovr/policy_net_1.py:86-94: flatten 400, per-sample standardise (unbiased std, NO eps), fc."""
    feat = pn1_unet(sd, torch.cat([image, context], dim=1), bf16).flatten(1)
    feat = (feat - feat.mean(dim=1, keepdim=True)) / feat.std(dim=1, keepdim=True)
    return F.linear(feat, sd["fc_final.weight"], sd["fc_final.bias"])


def gumbel_softmax_with_noise(logits, tau, expo):
    """F.gumbel_softmax(hard=False, dim=1) with the Exp(1) draw made explicit:
    gumbels = -log(expo); softmax((logits + gumbels) / tau). torch's implementation draws
    `torch.empty_like(logits).exponential_()` from the device's global generator."""
    return F.softmax((logits - expo.log()) / tau, dim=1)


def pn1_forward(sd, image, context, is_critic, expo=None, bf16=False):
    """rovr/policy_net_1.py:96-105. Actor: valid for b == 1 only (mean(dim=1) has no keepdim)."""
    logits = pn1_compute_logits(sd, image, context, bf16)
    if is_critic:
        return logits.squeeze(1)
    logits = (logits - logits.mean(dim=1)) / (logits.std(dim=(1,), keepdim=True) + 0.1)
    probs = gumbel_softmax_with_noise(logits, 0.5, expo)
    mx = probs.max(dim=1)
    return mx.indices, mx.values.log()


def pn1_logprob(sd, image, context, action, expo, bf16=False):
    """rovr/policy_net_1.py:107-114: gumbel-softmax on the UN-standardised logits, gather, log."""
    logits = pn1_compute_logits(sd, image, context, bf16)
    probs = gumbel_softmax_with_noise(logits, 0.5, expo)
    return probs.gather(1, action[:, None]).log().squeeze(1)


# ------------------------------------------------------------------------------------------------
# Policy network 2
# ------------------------------------------------------------------------------------------------
def pn2_video_conv(sd, image, bf16=False):
    """rovr/policy_net_2.py:41-60: 4 x [conv3x3 + BN + ReLU + pool] on [b,1,160,160] -> [b,1024]."""
    rt, wq = _storage(bf16)
    image = rt(image, False)

    def cbr(i, t):
        c, b = f"video_conv.{i}", f"video_conv.{i + 1}"
        raw = rt(F.conv2d(t, wq(sd[c + ".weight"]), sd[c + ".bias"], padding=1))
        return rt(F.relu(_bn_train(sd, b, raw)))

    t = F.max_pool2d(cbr(0, image), 8, 8)
    t = F.max_pool2d(cbr(4, t), 4, 4)
    t = F.max_pool2d(cbr(8, t), 1, 1)
    t = cbr(12, t)
    t = F.max_pool2d(t, 2, (2, 1))
    t = F.max_pool2d(t, 2, (2, 2))
    return t.flatten(1)


def pn2_final_fc(sd, v):
    """rovr/policy_net_2.py:63-69,71-79: five Linear layers with no activation between them."""
    for i in range(5):
        v = F.linear(v, sd[f"final_fc.{i}.weight"], sd[f"final_fc.{i}.bias"])
    return v


def pn2_masked_logits(sd, stacked, target):
    """rovr/policy_net_2.py:110-124: zero the target column(s), then (l - mean) / (std + .1) where
    mean(dim=1) has NO keepdim — for b == 20 it subtracts the mean of row j from column j;
    reproduced as written."""
    logits = pn2_final_fc(sd, stacked)
    idx = target.to(torch.int64).squeeze(1)
    logits = logits.scatter(1, idx, 0.0)
    return (logits - logits.mean(dim=1)) / (logits.std(dim=(1,), keepdim=True) + 0.1)


def pn2_forward(sd, image, context, target, is_critic, extra=None, expo=None, bf16=False):
    """rovr/policy_net_2.py:81-108."""
    if is_critic:
        image = image[:, None]
    stacked = torch.cat([pn2_video_conv(sd, image, bf16), context.squeeze(1)], dim=1)
    if extra is not None:
        return pn2_masked_logits(sd, stacked, target)
    if not is_critic:
        logits = pn2_masked_logits(sd, stacked, target)
        probs = gumbel_softmax_with_noise(logits, 0.7, expo)
        top = torch.topk(probs, k=2, dim=1)
        return top.indices, top.values.log().sum(1) / 2 + 0.69314
    mean = stacked.mean(dim=0, keepdim=True)
    std = stacked.std(dim=0, keepdim=True)
    return pn2_final_fc(sd, (stacked - mean) / (std + 0.001)).squeeze(1)


def pn2_logprob(sd, image, context, target, action, expo, bf16=False):
    """rovr/policy_net_2.py:127-141."""
    stacked = torch.cat([pn2_video_conv(sd, image[:, None], bf16), context.squeeze(1)], dim=1)
    logits = pn2_final_fc(sd, stacked).scatter(1, target.to(torch.int64), 0.0)
    probs = gumbel_softmax_with_noise(logits, 0.7, expo)
    pair = (probs[:, :, None] * probs[:, None, :]).flatten(1)
    flat_action = action[:, 0] * probs.shape[1] + action[:, 1]
    return pair.gather(1, flat_action[:, None]).log().sum(1) / 2 + 0.69314


def bn_running_update(x, running_mean, running_var, momentum=0.1):
    """What train-mode BatchNorm2d does to its buffers: unbiased variance, momentum 0.1."""
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=True)
    return (1 - momentum) * running_mean + momentum * mean, (1 - momentum) * running_var + momentum * var


# ------------------------------------------------------------------------------------------------
# common_layers blocks
# ------------------------------------------------------------------------------------------------
def layer_norm(sd, name, t):
    return F.layer_norm(t, (t.shape[-1],), sd[name + ".weight"], sd[name + ".bias"])


def mha(sd, name, q, kv, num_heads, drop_mask=None):
    """nn.MultiheadAttention(batch_first=True): packed in-proj [3E, E], per-head
    softmax(QK^T / sqrt(d)) V, out-proj. Returns the attention output only. drop_mask (optional,
    [B, heads, S, T], entries 0 or 1 / (1 - p)) multiplies the attention probabilities — the training-mode
    dropout of F.multi_head_attention_forward with the mask made explicit."""
    E = q.shape[-1]
    w, b = sd[name + ".in_proj_weight"], sd[name + ".in_proj_bias"]
    Q = F.linear(q, w[:E], b[:E])
    K = F.linear(kv, w[E:2 * E], b[E:2 * E])
    V = F.linear(kv, w[2 * E:], b[2 * E:])
    B, S, _ = Q.shape
    T = K.shape[1]
    d = E // num_heads
    Q = Q.view(B, S, num_heads, d).transpose(1, 2)
    K = K.view(B, T, num_heads, d).transpose(1, 2)
    V = V.view(B, T, num_heads, d).transpose(1, 2)
    att = torch.softmax(Q @ K.transpose(-1, -2) / math.sqrt(d), dim=-1)
    if drop_mask is not None:
        att = att * drop_mask
    o = (att @ V).transpose(1, 2).reshape(B, S, E)
    return F.linear(o, sd[name + ".out_proj.weight"], sd[name + ".out_proj.bias"])


def self_attention_block(sd, prefix, x, num_heads, drop_mask=None):
    """rovr/common_layers.py:54-64: x = LN(x); x = x + MHA(x, x, x) (residual on the normalised x)."""
    x = layer_norm(sd, prefix + "layer_norm", x)
    return x + mha(sd, prefix + "attention", x, x, num_heads, drop_mask)


def cross_attention_block(sd, prefix, x, enc, num_heads):
    """rovr/common_layers.py:66-78."""
    x = layer_norm(sd, prefix + "layer_norm", x)
    enc = layer_norm(sd, prefix + "layer_norm_encoder_output", enc)
    return x + mha(sd, prefix + "attention", x, enc, num_heads)


def feed_forward_block(sd, prefix, x, drop_mask=None):
    """rovr/common_layers.py:80-92: LN -> fc1 (E -> E/4) -> exact GELU -> dropout -> fc2. drop_mask
    (optional, shape of the hidden activation, entries 0 or 1 / (1 - p)) is the nn.Dropout mask."""
    x = layer_norm(sd, prefix + "layer_norm", x)
    h = F.gelu(F.linear(x, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"]))
    if drop_mask is not None:
        h = h * drop_mask
    return F.linear(h, sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"])


def encoder_block(sd, prefix, x, num_heads):
    """rovr/common_layers.py:94-104."""
    x = x + self_attention_block(sd, prefix + "attention.", x, num_heads)
    return x + feed_forward_block(sd, prefix + "feed_forward.", x)


def decoder_block(sd, prefix, x, enc, num_heads):
    """rovr/common_layers.py:106-118."""
    x = x + self_attention_block(sd, prefix + "attention.", x, num_heads)
    x = x + cross_attention_block(sd, prefix + "cross_attention.", x, enc, num_heads)
    return x + feed_forward_block(sd, prefix + "feed_forward.", x)


def image_positional_encoding(sd, prefix, x, num_image_patches):
    """rovr/common_layers.py:7-25: Linear(1, P^2 C) applied to arange(n^2), added to x."""
    pos = torch.arange(num_image_patches ** 2, dtype=torch.float32)[:, None]
    return x + F.linear(pos, sd[prefix + "positional_encoder.weight"], sd[prefix + "positional_encoder.bias"])[None]


def context_positional_encoding(sd, prefix, x, num_context_patches, num_context):
    """rovr/common_layers.py:27-52."""
    pp = torch.arange(num_context_patches ** 2, dtype=torch.float32)[:, None]
    cp = torch.arange(num_context, dtype=torch.float32)[:, None]
    patch = F.linear(pp, sd[prefix + "patch_positional_encoder.weight"], sd[prefix + "patch_positional_encoder.bias"])
    ctx = F.linear(cp, sd[prefix + "context_positional_encoder.weight"], sd[prefix + "context_positional_encoder.bias"])
    table = (patch[None, :, :] + ctx[:, None, :]).reshape(1, -1, patch.shape[-1])
    return x + table


# ------------------------------------------------------------------------------------------------
# ActionLSTM
# ------------------------------------------------------------------------------------------------
def action_lstm_step(sd, action, new_tensor, hx, cx):
    """rovr/action_lstm.py:19-38: LSTMCell(3 + 2304, hidden) + Linear(hidden, 19200) -> [b,3,80,80]."""
    inp = torch.cat([action.float() / 48, new_tensor.flatten(1)], dim=1)
    gates = F.linear(inp, sd["lstm.weight_ih"], sd["lstm.bias_ih"]) + F.linear(hx, sd["lstm.weight_hh"], sd["lstm.bias_hh"])
    i, f, g, o = gates.chunk(4, dim=1)
    cx = torch.sigmoid(f) * cx + torch.sigmoid(i) * torch.tanh(g)
    hx = torch.sigmoid(o) * torch.tanh(cx)
    out = F.linear(hx, sd["fc.weight"], sd["fc.bias"]).view(-1, 3, 80, 80)
    return out, hx, cx


# ------------------------------------------------------------------------------------------------
# bf16 storage emulation (to separate "inherent bf16 noise" from implementation error in tests)
# ------------------------------------------------------------------------------------------------
class _RoundTrip(torch.autograd.Function):
    """Identity whose forward value and backward gradient are rounded to bf16 and back — the
    storage precision of activations and activation-gradients on the B200 path."""

    @staticmethod
    def forward(ctx, t, round_grad):
        ctx.round_grad = round_grad
        return t.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return (g.to(torch.bfloat16).to(torch.float32) if ctx.round_grad else g), None


def localnet_step_bf16_storage(sd, x, context, target):
    """localnet_step with fp32 arithmetic but bf16 *storage* at the points where the B200 path
    stores bf16: packed input, conv/upconv weights (conv8 stays fp32), every post-ReLU
    activation, and every activation gradient. Differences between this and localnet_step are
    what any bf16 tensor-core implementation must show; differences between this and the CUDA
    path are implementation error (plus the occasional 1-ulp rounding flip)."""
    leaf = {k: (v.clone().requires_grad_(True) if k in LOCALNET_LIVE else v) for k, v in sd.items()}

    def w(name):
        p = leaf[name + ".weight"]
        return p + (p.to(torch.bfloat16).to(torch.float32) - p).detach()  # straight-through rounding

    rt = _RoundTrip.apply
    b = x.shape[0]
    h = torch.cat([x[:, None], context], dim=1).reshape(b, 9, x.shape[2], x.shape[3])
    h = rt(h, False)

    def conv(name, t):
        return rt(F.relu(F.conv2d(t, w(name), leaf[name + ".bias"], padding=1)), True)

    def up(name, t):
        return rt(F.relu(F.conv_transpose2d(t, w(name), leaf[name + ".bias"], stride=2)), True)

    def pool(t):
        return rt(F.max_pool2d(t, 2), True)

    e1 = conv("conv1", h)
    e2 = conv("conv2", pool(e1))
    e3 = conv("conv3", pool(e2))
    e4 = conv("conv4", pool(e3))
    d = conv("conv5", rt(torch.cat([up("upconv1", e4), e3], 1), True))
    d = conv("conv6", rt(torch.cat([up("upconv2", d), e2], 1), True))
    d = conv("conv7", rt(torch.cat([up("upconv3", d), e1], 1), True))
    y = torch.sigmoid(F.conv2d(d, leaf["conv8.weight"], leaf["conv8.bias"]))
    loss = F.mse_loss(y, target)
    loss.backward()
    return y.detach(), loss.detach(), {k: leaf[k].grad for k in LOCALNET_LIVE}


# ------------------------------------------------------------------------------------------------
# ResNet-50 frame-feature extractor (third-party arithmetic: torchvision.models.resnet50 and PIL via
# torchvision.transforms, exactly what rovr/resnet_extractor.py:8,18-23 calls)
# ------------------------------------------------------------------------------------------------
def resnet_randomise_bn(module, seed):
    """Deterministic non-trivial BatchNorm affine + running statistics, so that eval-mode folding is
    exercised and activations stay O(1) through the 16 residual blocks (default init has mean 0 /
    var 1 / gamma 1, which lets the residual stream grow by orders of magnitude)."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            c = m.num_features
            with torch.no_grad():
                m.weight.copy_(0.5 + 0.5 * torch.rand(c, generator=g))
                m.bias.copy_(0.2 * (torch.rand(c, generator=g) - 0.5))
                m.running_mean.copy_(0.2 * (torch.rand(c, generator=g) - 0.5))
                m.running_var.copy_(2.0 + 4.0 * torch.rand(c, generator=g))


def resnet_extractor_forward(resnet_seq, linear_w, linear_b, x):
    """rovr/resnet_extractor.py:25-47: per frame ToPILImage -> Resize((224,224)) -> ToTensor, the
    ResNet-50 trunk (children[:-1]) at batch 1, Linear(2048, 768) -> 3x16x16 tile pasted at
    (s // 5 * 16, s % 5 * 16) of a zero [b,3,80,80] map. resnet_seq must be in eval mode."""
    import torchvision.transforms as transforms
    prep = transforms.Compose([transforms.ToPILImage(), transforms.Resize((224, 224)), transforms.ToTensor()])
    b, s_len = x.shape[0], x.shape[1]
    fmap = torch.zeros((b, 3, 80, 80))
    for bi in range(b):
        for s in range(s_len):
            f = prep(x[bi, s]).unsqueeze(0)
            feat = resnet_seq(f)
            tile = F.linear(feat.view(-1), linear_w, linear_b).view(3, 16, 16)
            r, c = s // 5 * 16, s % 5 * 16
            fmap[bi, :, r:r + 16, c:c + 16] = tile
    return fmap


def video_processor_forward(resnet_seq, linear_w, linear_b, x):
    """`VideoProcessor` (rovr/rovr.py:61,107; rovr/imitation_learning.py:58,78) — the file is ABSENT from the
    reference: PARITY UNPINNED AND UN-ORACLED by it. Restated from its call sites as the 32-pixel-tile sibling
    of rovr/resnet_extractor.py:25-47: per frame ToPILImage -> Resize((224,224)) -> ToTensor, eval-mode ResNet-50
    trunk at batch 1, Linear(2048, 1024) -> flattened_frames[b, S, 1024]; each vector as a 1x32x32 tile pasted at
    (s // 5 * 32, s % 5 * 32) of a zero [b, 1, 160, 160] map. Returns (encoded_frames, flattened_frames)."""
    import torchvision.transforms as transforms
    prep = transforms.Compose([transforms.ToPILImage(), transforms.Resize((224, 224)), transforms.ToTensor()])
    b, s_len = x.shape[0], x.shape[1]
    enc = torch.zeros((b, 1, 160, 160))
    flat = []
    for bi in range(b):
        for s in range(s_len):
            feat = resnet_seq(prep(x[bi, s]).unsqueeze(0))
            vec = F.linear(feat.view(-1), linear_w, linear_b)
            flat.append(vec)
            r, c = s // 5 * 32, s % 5 * 32
            enc[bi, :, r:r + 32, c:c + 32] = vec.view(1, 32, 32)
    return enc, torch.stack(flat).view(b, s_len, -1)


# ------------------------------------------------------------------------------------------------
# LPIPS (net='vgg') perceptual loss — SURVEY §8f-1. Call sites: rovr/train_local_net_unet.py:91,109
# (`lpips.LPIPS(net='vgg')(y_hat, target).mean()`, normalize=False) and rovr/rovr.py:54,84,255
# (normalize=True). The `lpips` package (version unpinned by the reference) is NOT in this image and
# its trained weights need a download: PARITY UNPINNED against the package. This is a restatement of
# its published v0.1 algorithm (Zhang et al. 2018, `lpips/lpips.py`):
#     ScalingLayer: (x - shift) / scale, shift = [-.030, -.088, -.188], scale = [.458, .448, .450]
#     torchvision VGG16 `features` tapped after relu1_2, relu2_2, relu3_3, relu4_3, relu5_3
#     normalize_tensor: x / (sqrt(sum_c x^2) + 1e-10)
#     per tap: 1x1 conv (no bias, lin weights [1, C, 1, 1]) of (f0 - f1)^2, spatial mean; sum over taps
#     (the package runs in eval mode: its Dropout in front of the 1x1 conv is inactive)
# The VGG16 structure is pinned to torchvision.models.vgg16 (tests/golden/lpips.npz is generated by
# running torchvision's own `features` with the seeded weights below).
# ------------------------------------------------------------------------------------------------
LPIPS_SHIFT = (-0.030, -0.088, -0.188)
LPIPS_SCALE = (0.458, 0.448, 0.450)
# torchvision vgg16().features indices of the 13 convolutions, grouped by LPIPS slice; a MaxPool2d(2, 2)
# precedes every slice but the first
VGG16_SLICES = [[(0, 3, 64), (2, 64, 64)], [(5, 64, 128), (7, 128, 128)],
                [(10, 128, 256), (12, 256, 256), (14, 256, 256)],
                [(17, 256, 512), (19, 512, 512), (21, 512, 512)],
                [(24, 512, 512), (26, 512, 512), (28, 512, 512)]]
LPIPS_CHNS = [64, 128, 256, 512, 512]


def lpips_state_dict(seed=0):
    """Deterministic weights in the key layout of lpips.LPIPS(net='vgg').state_dict():
    net.slice{k}.{idx}.weight/bias (torchvision feature indices), lin{k}.model.1.weight [1, C, 1, 1]
    (non-negative, like the trained ones), scaling_layer.shift / scale."""
    gen = torch.Generator().manual_seed(seed)
    sd = {"scaling_layer.shift": torch.tensor(LPIPS_SHIFT).view(1, 3, 1, 1),
          "scaling_layer.scale": torch.tensor(LPIPS_SCALE).view(1, 3, 1, 1)}
    for k, convs in enumerate(VGG16_SLICES):
        for idx, cin, cout in convs:
            sd[f"net.slice{k + 1}.{idx}.weight"] = _fill(gen, (cout, cin, 3, 3), math.sqrt(6.0 / (cin * 9)))
            sd[f"net.slice{k + 1}.{idx}.bias"] = _fill(gen, (cout,), 0.1)
    for k, c in enumerate(LPIPS_CHNS):
        sd[f"lin{k}.model.1.weight"] = (_fill(gen, (1, c, 1, 1), 1.0).abs() + 0.05) * (4.0 / c)
    return sd


def lpips_vgg_features(sd, x, bf16=False):
    """The five tapped activations of torchvision VGG16 `features` for an already-scaled input.
    bf16=True emulates the storage points of the B200 path (fp32 arithmetic; the packed input, the conv
    weights, every post-ReLU activation and every activation gradient rounded to bf16)."""
    rt, wq = _storage(bf16)
    x = rt(x)
    feats = []
    for k, convs in enumerate(VGG16_SLICES):
        if k > 0:
            x = F.max_pool2d(x, 2, 2)
        for idx, _, _ in convs:
            x = rt(F.relu(F.conv2d(x, wq(sd[f"net.slice{k + 1}.{idx}.weight"]), sd[f"net.slice{k + 1}.{idx}.bias"], padding=1)))
        feats.append(x)
    return feats


def lpips_vgg(sd, in0, in1, normalize=False, bf16=False):
    """lpips.LPIPS(net='vgg').forward(in0, in1, normalize=normalize) -> [N, 1, 1, 1]."""
    if normalize:
        in0, in1 = 2 * in0 - 1, 2 * in1 - 1
    shift, scale = sd["scaling_layer.shift"], sd["scaling_layer.scale"]
    f0s = lpips_vgg_features(sd, (in0 - shift) / scale, bf16)
    f1s = lpips_vgg_features(sd, (in1 - shift) / scale, bf16)
    val = 0.0
    for k, (f0, f1) in enumerate(zip(f0s, f1s)):
        n0 = f0 / (torch.sqrt(torch.sum(f0 ** 2, dim=1, keepdim=True)) + 1e-10)
        n1 = f1 / (torch.sqrt(torch.sum(f1 ** 2, dim=1, keepdim=True)) + 1e-10)
        d = F.conv2d((n0 - n1) ** 2, sd[f"lin{k}.model.1.weight"])
        val = val + d.mean([2, 3], keepdim=True)
    return val
