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
  * frame indices stay on the device (index_select) — no host synchronisation inside the rollout;
  * `graphed=True`: a time-step (actor forward with its gumbel draw, context gather, LocalNet forward,
    LPIPS reward, re-encode of the reconstructed frame, tile paste, bookkeeping) is ONE CUDA graph replayed
    S times with a device-side step counter, and each PPO update is two graph replays (critic and actor
    forward + backward) around the eager Adam steps: an RL iteration is ~4200 kernel launches of a few
    microseconds each, i.e. bound by launch overhead when run eagerly.
"""
import torch

import ops


class ROVRStep:
    def __init__(self, actor2, critic2, local_net, lpips_fn, video_processor, actor_optimizer=None,
                 critic_optimizer=None, clip=0.2, n_updates_per_ppo=5, averager=None, graphed=False):
        self.actor2, self.critic2, self.local_net = actor2, critic2, local_net
        self.lpips, self.video_processor = lpips_fn, video_processor
        self.actor_optimizer = actor_optimizer or torch.optim.Adam(actor2.parameters(), lr=2e-4)      # rovr/rovr.py:58-59
        self.critic_optimizer = critic_optimizer or torch.optim.Adam(critic2.parameters(), lr=2e-4)
        self.clip = clip
        self.num_updates_per_ppo = n_updates_per_ppo
        self.averagers = averager            # optional (actor, critic) data_parallel.GradientAverager pair
        self.graphed = graphed
        self.replayed_launches = 0           # kernels of librovr_b200 launched through graph replays (bench.py reports them)
        self.launches_per_time_step = 0
        self._ts_graph = None                # (key, graph, state) of the captured time-step
        self._ppo_graphs = None

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

    # -- the same rollout with ONE CUDA graph per time-step --------------------------------------------------
    def _time_step(self, st):
        """One time-step on static state `st`; every index is a device tensor (st["j"] = current step)."""
        K, S = st["K"], st["S"]
        video, org, flat_v = st["video"], st["org"], st["flat_v"]
        c, h, w = video.shape[2:]
        j = st["j"]                                                           # int64 [1]
        tgt = j.view(1, 1, 1)
        tf_all = st["flattened"].index_select(1, j)                           # [K, 1, 1024]
        ctx_idx = []
        for k in range(K):
            idx, logp = self.actor2(st["encoded"][k:k + 1], tf_all[k:k + 1], tgt)
            st["obs_enc"][k].index_copy_(0, j, st["encoded"][k])              # [S, 160, 160] <- [1, 160, 160]
            st["acs"][k].index_copy_(0, j, idx)
            st["logp"][k].index_copy_(0, j, logp.view(1, 1))
            ctx_idx.append(idx[0])
        ctx_idx = torch.stack(ctx_idx)
        gather = (st["karange"][:, None] * S + ctx_idx).reshape(-1)
        context = flat_v.index_select(0, gather).view(K, 2, c, h, w)
        frame_j = video.index_select(1, j)[:, 0]
        y_hat = self.local_net(frame_j, context)
        reward = self.lpips(y_hat, org.index_select(1, j)[:, 0], normalize=True).view(K)
        st["recon"].index_copy_(1, j, y_hat[:, None])
        feats = self.video_processor._encode_batch(y_hat)
        ops.mosaic_paste(feats.contiguous(), st["encoded"], batch=st["karange"], slot=j.expand(K).contiguous(),
                         tile=self.video_processor.TILE)
        cur = st["curr_loss"].index_select(1, j)[:, 0]
        st["rewards"].index_copy_(1, j, (-(reward - cur))[:, None])
        st["curr_loss"].index_copy_(1, j, reward[:, None])
        j.add_(1)

    @torch.no_grad()
    def rollout_graphed(self, video, org_video):
        K, S, c, h, w = video.shape
        dev = video.device
        key = (K, S, c, h, w, dev)
        if self._ts_graph is None or self._ts_graph[0] != key:
            st = {"K": K, "S": S, "video": video.float().clone(), "org": org_video.float().clone(),
                  "j": torch.zeros(1, dtype=torch.int64, device=dev), "karange": torch.arange(K, device=dev),
                  "encoded": torch.zeros((K, 1, 160, 160), device=dev), "flattened": torch.zeros((K, S, 1024), device=dev),
                  "obs_enc": [torch.zeros((S, 160, 160), device=dev) for _ in range(K)],
                  "acs": [torch.zeros((S, 2), dtype=torch.int64, device=dev) for _ in range(K)],
                  "logp": [torch.zeros((S, 1), device=dev) for _ in range(K)],
                  "recon": torch.zeros((K, S, c, h, w), device=dev), "rewards": torch.zeros((K, S), device=dev),
                  "curr_loss": torch.zeros((K, S), device=dev)}
            st["flat_v"] = st["video"].view(K * S, c, h, w)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                                     # warm-up: packs weights, sizes workspaces
                self._time_step(st)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            import _native
            n0 = _native.lib.rovr_launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                self._time_step(st)
            self.launches_per_time_step = int(_native.lib.rovr_launch_count() - n0)   # kernels of this library per replay
            self._ts_graph = (key, graph, st)
        _, graph, st = self._ts_graph
        st["video"].copy_(video)
        st["org"].copy_(org_video)
        flat_o = st["org"].view(K * S, c, h, w)
        st["curr_loss"].copy_(self.lpips(st["flat_v"], flat_o, normalize=True).view(K, S))
        enc, flat = self.video_processor(st["video"])
        st["encoded"].copy_(enc)
        st["flattened"].copy_(flat)
        st["recon"].copy_(st["video"])
        st["j"].zero_()
        for _ in range(S):
            graph.replay()
        self.replayed_launches += S * self.launches_per_time_step
        rtg = torch.flip(torch.cumsum(torch.flip(st["rewards"], [1]), 1), [1])
        steps = torch.arange(S, dtype=torch.int64, device=dev).unsqueeze(-1)
        out = [((st["obs_enc"][k].clone(), st["flattened"][k].clone(), steps), st["acs"][k].clone(), st["logp"][k].clone(),
                rtg[k].view(-1, 1)) for k in range(K)]
        return out, st["recon"].clone()

    def _ppo_graphed(self, info, device):
        from graphs import GraphedFunction
        obs, acs, log_prob, rtgs = info
        actor, critic = self.actor2, self.critic2
        key = tuple(tuple(t.shape) for t in (*obs, acs, log_prob, rtgs))
        if self._ppo_graphs is not None and self._ppo_graphs[2] != key:      # another clip length: capture again
            for g in self._ppo_graphs[:2]:
                g.close()
            self._ppo_graphs = None
        if self._ppo_graphs is None:
            s_obs = tuple(t.clone() for t in obs)
            s_acs, s_lp, s_rtg = acs.clone(), log_prob.clone(), rtgs.clone()
            s_A = torch.zeros((rtgs.shape[0], rtgs.shape[0]), device=rtgs.device)

            def critic_fn(o0, o1, o2, rtg):
                V = critic(o0, o1, o2, device)
                loss = torch.nn.functional.mse_loss(V, rtg.squeeze(1))
                loss.backward()
                return loss.detach(), V.detach()

            def actor_fn(o0, o1, o2, a, lp, A):
                cur = actor.logprob(o0, o1, o2, a, A.device).unsqueeze(1)
                ratio = torch.exp(cur - lp)
                loss = -torch.min(ratio * A, torch.clamp(ratio, 1 - self.clip, 1 + self.clip) * A).mean()
                loss.backward()
                return loss.detach()
            gc = GraphedFunction(critic_fn, (*s_obs, s_rtg), modules=[critic])
            ga = GraphedFunction(actor_fn, (*s_obs, s_acs, s_lp, s_A), modules=[actor])
            self._ppo_graphs = (gc, ga, key)
        gc, ga = self._ppo_graphs[:2]
        with torch.no_grad():
            V = critic(*obs, device)
        A_k = rtgs - V.detach()
        A_k = (A_k - A_k.mean()) / (A_k.std() + 1e-10)
        losses = []
        for _ in range(self.num_updates_per_ppo):
            # GraphedFunction leaves this replay's gradients in the parameters' static .grad tensors
            c_loss, _ = gc(*obs, rtgs)
            if self.averagers is not None:
                self.averagers[1].average()
            self.critic_optimizer.step()
            a_loss = ga(*obs, acs, log_prob, A_k)
            self.replayed_launches += gc.launches + ga.launches
            if self.averagers is not None:
                self.averagers[0].average()
            self.actor_optimizer.step()
            losses.append((a_loss.clone(), c_loss.clone()))
        return losses

    # -- PPO (rovr/rovr.py:281-337) ------------------------------------------------------------------------
    def ppo(self, info, device=None):
        if self.graphed:
            return self._ppo_graphed(info, device)
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
        infos, recon = (self.rollout_graphed if self.graphed else self.rollout)(video, org_video)
        return [self.ppo(info, video.device) for info in infos], recon
