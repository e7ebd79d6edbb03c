"""The RL step of rovr/rovr.py restated on the drop-ins (rovr_step.ROVRStep) against the
same step restated on the CPU oracle with the SAME gumbel noise: the context-frame indices the policy
selects must be bit-exact at every one of the 20 time-steps (so both trajectories stay identical), rewards /
rewards-to-go within the bf16 tolerance, PPO losses within tolerance; and batching K clips must not change
a clip's trajectory."""
import copy
import warnings

import pytest
import torch
import torch.nn.functional as F

import rovr_oracle as O

pytestmark = pytest.mark.gpu
S, HW = 20, 64


def _dev():
    import _native
    _native.require_device()
    return torch.device("cuda:0")


def _clips(k, seed):
    g = torch.Generator().manual_seed(seed)
    org = torch.rand((k, S, 3, 8, 8), generator=g)
    org = F.interpolate(org.view(k * S, 3, 8, 8), size=(HW, HW), mode="bilinear", align_corners=False).view(k, S, 3, HW, HW)
    vid = org.clone()
    for n in range(S):                                           # raster-scan box, like rovr/video_ds.py:62-87
        x0, y0 = (n % 4) * 12, (n // 4) * 10
        vid[:, n, :, y0:y0 + 24, x0:x0 + 32] = 0
    return vid.contiguous(), org.contiguous()


class _Noise:
    """Deterministic Exp(1) draws shared by the module (monkeypatched exponential_like) and the oracle."""

    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.log = []

    def __call__(self, logits):
        e = torch.empty(tuple(logits.shape)).exponential_(generator=self.g)
        self.log.append(e)
        return e.to(logits.device)


def _build(dev):
    from lpips_vgg import LPIPS
    from local_net import LocalNetworkUNetNorm
    from policy_net_2 import PolicyNetwork2UNet
    from video_processor import VideoProcessor
    warnings.filterwarnings("ignore")
    torch.manual_seed(11)
    vp = VideoProcessor()
    O.resnet_randomise_bn(vp.resnet, 61)
    parts = {"sd_actor": O.pn2_state_dict(0, False), "sd_critic": O.pn2_state_dict(1, True), "sd_local": O.localnet_state_dict(0),
             "sd_lpips": O.lpips_state_dict(0), "resnet": copy.deepcopy(vp.resnet).eval(),
             "lw": vp.linear.weight.detach().clone(), "lb": vp.linear.bias.detach().clone()}
    actor, critic = PolicyNetwork2UNet(), PolicyNetwork2UNet(is_critic=True)
    actor.load_state_dict(parts["sd_actor"]); critic.load_state_dict(parts["sd_critic"])
    local = LocalNetworkUNetNorm(freeze=True)                    # rovr/rovr.py:37
    local.load_state_dict(parts["sd_local"])
    lp = LPIPS(net="vgg")
    lp.load_state_dict(parts["sd_lpips"])
    mods = [m.to(dev) for m in (actor.train(), critic.train(), local, lp, vp)]
    return mods, parts


def _oracle_rollout(parts, vid, org, noise):
    """rovr/rovr.py:81-209 for one clip on the CPU oracle (same deviations as rovr_step.ROVRStep)."""
    sd_a, sd_l, sd_p = parts["sd_actor"], parts["sd_local"], parts["sd_lpips"]
    sd_a = {k: v.clone() for k, v in sd_a.items()}
    with torch.no_grad():
        curr = O.lpips_vgg(sd_p, vid[0], org[0], normalize=True).view(S).clone()
        enc, flat = O.video_processor_forward(parts["resnet"], parts["lw"], parts["lb"], vid)
        idxs, logps, rewards, encs = [], [], [], []
        for j in range(S):
            tf = flat[:, j:j + 1, :]
            encs.append(enc[0, 0].clone())
            expo = noise(torch.empty(1, 20))
            idx, logp = O.pn2_forward(sd_a, enc, tf, torch.full((1, 1, 1), j), False, expo=expo)
            O_bn = None  # (BatchNorm running statistics do not enter the train-mode forward)
            ctx = vid[:, idx[0]]                                 # [1, 2, 3, H, W]
            y_hat = O.localnet_forward(sd_l, vid[:, j], ctx)
            reward = O.lpips_vgg(sd_p, y_hat, org[:, j], normalize=True).view(())
            e1, _ = O.video_processor_forward(parts["resnet"], parts["lw"], parts["lb"], y_hat[:, None])
            r, c = j // 5 * 32, j % 5 * 32
            enc = enc.clone()
            enc[0, :, r:r + 32, c:c + 32] = e1[0, :, 0:32, 0:32]
            rewards.append(-(reward - curr[j]))
            curr[j] = reward
            idxs.append(idx[0]); logps.append(logp)
        rewards = torch.stack(rewards)
        rtg = torch.flip(torch.cumsum(torch.flip(rewards, [0]), 0), [0]).view(-1, 1)
    obs = (torch.stack(encs), flat[0], torch.arange(S).unsqueeze(-1))
    return obs, torch.stack(idxs), torch.stack(logps), rtg


def test_rollout_indices_bit_exact_and_ppo_vs_oracle(monkeypatch):
    import policy_net_2 as M
    from rovr_step import ROVRStep
    dev = _dev()
    (actor, critic, local, lp, vp), parts = _build(dev)
    vid, org = _clips(1, 5)
    noise = _Noise(123)
    monkeypatch.setattr(M, "exponential_like", noise)
    step = ROVRStep(actor, critic, local, lp, vp, n_updates_per_ppo=2)
    infos, recon = step.rollout(vid.to(dev), org.to(dev))
    (obs, acs, logp, rtg), = infos
    assert obs[0].shape == (S, 160, 160) and obs[1].shape == (S, 1024) and obs[2].shape == (S, 1)
    assert acs.shape == (S, 2) and acs.dtype == torch.int64 and logp.shape == (S, 1) and rtg.shape == (S, 1)
    o_obs, o_acs, o_logp, o_rtg = _oracle_rollout(parts, vid, org, _Noise(123))
    print("selected context frames:", acs.cpu().tolist())
    assert torch.equal(acs.cpu(), o_acs), "context-frame indices differ from the oracle rollout"
    assert (acs.cpu() != torch.arange(S)[:, None]).all(), "the target frame itself must be masked out"
    rel = lambda a, b: ((a.detach().float().cpu() - b).norm() / (b.norm() + 1e-20)).item()
    print(f"rollout vs oracle: mosaics {rel(obs[0], o_obs[0]):.3e} features {rel(obs[1], o_obs[1]):.3e} "
          f"log-prob {rel(logp, o_logp):.3e} rewards-to-go {rel(rtg, o_rtg):.3e}")
    assert rel(obs[0], o_obs[0]) < 2e-2 and rel(obs[1], o_obs[1]) < 2e-2
    assert rel(logp, o_logp) < 2e-2 and rel(rtg, o_rtg) < 5e-2
    # PPO (rovr/rovr.py:281-337): two updates vs the same arithmetic on the oracle with torch.optim.Adam
    losses = step.ppo((obs, acs, logp, rtg), dev)
    sd_a = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
            for k, v in parts["sd_actor"].items()}
    sd_c = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
            for k, v in parts["sd_critic"].items()}
    opt_a = torch.optim.Adam([v for v in sd_a.values() if v.requires_grad], lr=2e-4)
    opt_c = torch.optim.Adam([v for v in sd_c.values() if v.requires_grad], lr=2e-4)
    ob = tuple(t.detach().cpu() for t in obs)
    acs_c, logp_c, rtg_c = acs.cpu(), logp.detach().cpu(), rtg.detach().cpu()
    onoise = _Noise(123)
    onoise.g.set_state(noise.g.get_state()) if False else None
    # the module drew S actor draws during the rollout, then one [S, 20] draw per logprob call
    replay = iter(noise.log[S:])
    V = O.pn2_forward(sd_c, ob[0], ob[1], ob[2], True)
    A = rtg_c - V.detach()
    A = (A - A.mean()) / (A.std() + 1e-10)
    for u in range(2):
        V = O.pn2_forward(sd_c, ob[0], ob[1], ob[2], True)
        clp = O.pn2_logprob(sd_a, ob[0], ob[1], ob[2], acs_c, next(replay)).unsqueeze(1)
        ratio = torch.exp(clp - logp_c)
        al = -torch.min(ratio * A, torch.clamp(ratio, 0.8, 1.2) * A).mean()
        cl = F.mse_loss(V, rtg_c.squeeze(1))
        opt_c.zero_grad(); cl.backward(); opt_c.step()
        opt_a.zero_grad(); al.backward(); opt_a.step()
        print(f"ppo update {u}: actor loss {float(losses[u][0]):.5f} vs {float(al):.5f}, critic loss "
              f"{float(losses[u][1]):.5f} vs {float(cl):.5f}")
        # update 0 sees identical weights: north_star tolerance. Update 1 comes after an Adam step of both nets, and the
        # critic divides every feature by (its std over the 20 rows + 1e-3) (rovr/policy_net_2.py:104-106): 1e-3-level
        # differences in the first step's gradients are amplified (its loss jumps 1 -> 32): 1e-1 there.
        tol = 2e-2 if u == 0 else 1e-1
        assert abs(float(losses[u][0]) - float(al)) < tol * max(abs(float(al)), 1e-2) + 2e-3
        assert abs(float(losses[u][1]) - float(cl)) < tol * max(abs(float(cl)), 1e-2)


def test_batched_clips_keep_each_trajectory(monkeypatch):
    import policy_net_2 as M
    from rovr_step import ROVRStep
    dev = _dev()
    (actor, critic, local, lp, vp), _ = _build(dev)
    vid, org = _clips(2, 6)
    step = ROVRStep(actor, critic, local, lp, vp)
    # one noise row per (time-step, clip): K = 2 consumes rows in (j, k) order; the single-clip runs replay them
    master = _Noise(77)
    monkeypatch.setattr(M, "exponential_like", master)
    infos2, _ = step.rollout(vid.to(dev), org.to(dev))
    rows = list(master.log)
    for k in range(2):
        it = iter(rows[k::2])
        monkeypatch.setattr(M, "exponential_like", lambda logits, it=it: next(it).to(logits.device))
        (single,), _ = step.rollout(vid[k:k + 1].to(dev), org[k:k + 1].to(dev))
        assert torch.equal(single[1], infos2[k][1]), f"clip {k}: batching changed the selected frames"
        assert torch.allclose(single[3], infos2[k][3], rtol=1e-3, atol=1e-5)


def test_graphed_rollout_matches_eager():
    """graphed=True: one CUDA graph per time-step (replayed 20 times with a device-side step counter) and two per
    PPO update give the eager trajectory — same device RNG stream, so the selected frames are identical."""
    from rovr_step import ROVRStep
    dev = _dev()
    (actor, critic, local, lp, vp), parts = _build(dev)
    vid, org = _clips(1, 9)
    eager = ROVRStep(actor, critic, local, lp, vp, n_updates_per_ppo=1)
    graphed = ROVRStep(actor, critic, local, lp, vp, n_updates_per_ppo=1, graphed=True)
    # capture first (warm-up + capture consume random numbers and update BatchNorm running statistics — neither
    # enters the train-mode forward values), then run both from the same generator state
    graphed.rollout_graphed(vid.to(dev), org.to(dev))
    torch.manual_seed(1234)
    (e_info,), e_rec = eager.rollout(vid.to(dev), org.to(dev))
    torch.manual_seed(1234)
    (g_info,), g_rec = graphed.rollout_graphed(vid.to(dev), org.to(dev))
    assert torch.equal(e_info[1], g_info[1]), (e_info[1].tolist(), g_info[1].tolist())
    assert torch.allclose(e_info[3], g_info[3], rtol=1e-4, atol=1e-6)
    assert torch.allclose(e_info[2], g_info[2], rtol=1e-4, atol=1e-6)
    assert torch.allclose(e_info[0][0], g_info[0][0], rtol=1e-4, atol=1e-6) and torch.equal(e_info[0][2], g_info[0][2])
    assert torch.allclose(e_rec, g_rec, rtol=1e-4, atol=1e-6)
    # a PPO update through the graphs changes both networks and returns finite losses
    w0 = actor.final_fc[4].weight.detach().clone()
    c0 = critic.final_fc[4].weight.detach().clone()
    losses = graphed.ppo(g_info, dev)
    assert len(losses) == 1 and all(torch.isfinite(v) for v in losses[0])
    assert not torch.equal(actor.final_fc[4].weight, w0) and not torch.equal(critic.final_fc[4].weight, c0)
