"""The reinforcement-learning step of rovr/rovr.py (rollout :81-249, train_local_network :252-265,
compute_rewards_to_go :268-278, ppo :281-337) RESTATED on the B200 drop-in modules — SURVEY §8f-3,
BASELINE.json configs[3]. `rovr.py` itself cannot be imported (it needs lpips, matplotlib, the
absent video_processor.py and checkpoint files, and trains at import time), so this is the step
logic only, without TensorBoard, the matplotlib displays and the RAFT optical-flow diagnostics
(:340-367, logged but never part of the reward: `rewards[-1] = rewards[-1] - spatio_loss` is
commented out, :229).

Differences from the reference, all deliberate:
  * clips are BATCHED: `rollout` takes K clips [K, S, 3, H, W]; LocalNet, LPIPS and the frame
    encoder run once per time-step on all K clips (the reference hard-codes batch 1, rovr/test.py:18).
    The PolicyNetwork2UNet actor is called per clip (its `forward` is only shape-valid for b == 1,
    rovr/policy_net_2.py:122), and PPO runs per clip, so BatchNorm statistics and the critic's
    batch-dim standardisation see exactly the reference's per-clip batch of S rows;
  * the actor's third argument is the target frame index `j` — what the reference RECORDS as the
    PPO observation (`torch.tensor(j)`, :147) — instead of `torch.tensor(target_frame)` (:141), which
    passes the 1024-d feature vector where `scatter_` needs frame indices;
  * the "exp_" branch (:163-175, a second no-grad LocalNet pass on the previous two frames, used for
    a visualisation only) is not run;
  * frame indices stay on the device (index_select) — no host synchronisation inside the rollout.
"""
import torch


class ROVRStep:
    def __init__(self, actor2, critic2, local_net, lpips_fn, video_processor, actor_optimizer=None,
                 critic_optimizer=None, clip=0.2, n_updates_per_ppo=5, averager=None):
        self.actor2, self.critic2, self.local_net = actor2, critic2, local_net
        self.lpips, self.video_processor = lpips_fn, video_processor
        self.actor_optimizer = actor_optimizer or torch.optim.Adam(actor2.parameters(), lr=2e-4)      # rovr/rovr.py:58-59
        self.critic_optimizer = critic_optimizer or torch.optim.Adam(critic2.parameters(), lr=2e-4)
        self.clip = clip
        self.num_updates_per_ppo = n_updates_per_ppo
        self.averagers = averager            # optional (actor, critic) data_parallel.GradientAverager pair

    # -- rollout (rovr/rovr.py:81-249) ---------------------------------------------------------------------
    @torch.no_grad()
    def rollout(self, video, org_video):
        """video / org_video: [K, S, 3, H, W] corrupted / clean clips in [0, 1]. Returns a list of K
        (obs, acs, log_prob, rtg) tuples shaped like the reference's, plus the reconstructed clips."""
        K, S, c, h, w = video.shape
        dev = video.device
        flat_v, flat_o = video.reshape(K * S, c, h, w), org_video.reshape(K * S, c, h, w)
        curr_loss = self.lpips(flat_v.float(), flat_o.float(), normalize=True).view(K, S).clone()      # :84
        encoded, flattened = self.video_processor(video)                 # [K,1,160,160], [K,S,1024]  (:107)
        recon = video.clone()
        obs = [([], [], []) for _ in range(K)]
        acs = [[] for _ in range(K)]
        logps = [[] for _ in range(K)]
        rewards = []
        karange = torch.arange(K, device=dev)
        for j in range(S):                                               # time_steps == vid_length (rovr/test.py:13-14)
            tgt = torch.full((1, 1, 1), j, dtype=torch.int64, device=dev)
            ctx_idx = []
            for k in range(K):
                tf = flattened[k:k + 1, j:j + 1, :]                      # [1, 1, 1024]  (:132)
                idx, logp = self.actor2(encoded[k:k + 1].float(), tf.float(), tgt)      # :141
                obs[k][0].append(encoded[k, 0].clone())
                obs[k][1].append(tf[0, 0])
                obs[k][2].append(j)
                acs[k].append(idx[0])
                logps[k].append(logp)
                ctx_idx.append(idx[0])
            ctx_idx = torch.stack(ctx_idx)                               # [K, 2] frame indices, on the device
            gather = (karange[:, None] * S + ctx_idx).reshape(-1)
            context = flat_v.index_select(0, gather).view(K, 2, c, h, w)                   # :153-162
            y_hat = self.local_net(video[:, j].float(), context.float())                    # :253
            reward = self.lpips(y_hat, org_video[:, j].float(), normalize=True).view(K)     # :255, :183
            recon[:, j] = y_hat
            encoded = self.video_processor.insert_encoded_frame_batch(
                torch.full((K, 1), j, dtype=torch.int64), y_hat, encoded)                   # :200
            rewards.append(-(reward - curr_loss[:, j]))                                     # :202
            curr_loss[:, j] = reward                                                        # :205
        rewards = torch.stack(rewards, dim=1)                                               # [K, S]
        rtg = torch.flip(torch.cumsum(torch.flip(rewards, [1]), 1), [1])                    # :268-278, gamma = 1
        out = []
        for k in range(K):
            o = (torch.stack(obs[k][0]), torch.stack(obs[k][1]),
                 torch.tensor(obs[k][2], dtype=torch.int64, device=dev).unsqueeze(-1))      # :213
            out.append((o, torch.stack(acs[k]), torch.stack(logps[k]), rtg[k].view(-1, 1)))
        return out, recon

    # -- PPO (rovr/rovr.py:281-337) ------------------------------------------------------------------------
    def ppo(self, info, device=None):
        obs, acs, log_prob, rtgs = info
        actor, critic = self.actor2, self.critic2
        with torch.no_grad():
            V = critic(*obs, device)
        A_k = rtgs - V.detach()                                            # :303 ([S,1] - [S] broadcasts to [S,S], as written)
        A_k = (A_k - A_k.mean()) / (A_k.std() + 1e-10)
        losses = []
        for _ in range(self.num_updates_per_ppo):
            V = critic(*obs, device)
            curr_log_prob = actor.logprob(*obs, acs, A_k.device).unsqueeze(1)
            ratio = torch.exp(curr_log_prob - log_prob)
            L1 = ratio * A_k
            L2 = torch.clamp(ratio, 1 - self.clip, 1 + self.clip) * A_k
            actor_loss = -torch.min(L1, L2).mean()
            critic_loss = torch.nn.functional.mse_loss(V, rtgs.squeeze(1))
            self.critic_optimizer.zero_grad()
            critic_loss.backward()
            if self.averagers is not None:
                self.averagers[1].average()
            self.critic_optimizer.step()
            self.actor_optimizer.zero_grad()
            actor_loss.backward()
            if self.averagers is not None:
                self.averagers[0].average()
            self.actor_optimizer.step()
            losses.append((actor_loss.detach(), critic_loss.detach()))
        return losses

    def train(self, video, org_video):
        """rovr/rovr.py:68-79: one rollout of every clip, then PPO on PolicyNetwork2UNet per clip."""
        infos, recon = self.rollout(video, org_video)
        return [self.ppo(info, video.device) for info in infos], recon
