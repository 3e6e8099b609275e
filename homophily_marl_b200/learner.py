"""Homophily IQL learner step for a sharded learner.

Same constructor and public methods as ``HomophilyLearner`` (src/learners/homophily_learner.py:11-49, 249-288) --
``train(batch, t_env, episode_num)``, ``cuda()``, ``save_models(path)``, ``load_models(path)`` -- so that ``run_sequential``
(src/run.py:132-135, 181-212) drives it as ``learners.REGISTRY['homophily_learner_b200']``.  What differs:

* the incentive bookkeeping of lines 98-115 is ONE kernel (``ssd_incentive``, incentive.py) instead of ~20 torch launches over
  int64 [bs, t, n, n, 1] tensors;
* the similarity clusters of lines 184-203 are computed on the device.  The reference runs pyclustering's x-means (k-means++
  seeded, on the CPU) over rows of ``[rewards_t, clean_num_t]`` in {0,1}^2 and only ever asks whether two rows fell into the
  same cluster; here a cluster IS a distinct point (PARITY UNPINNED against pyclustering, which is neither vendored nor
  deterministic -- identical to the stand-in in baseline/stubs that the reference learner runs with in this repo's tests);
* data parallelism: every rank trains on its shard of the sampled episodes and the gradients of BOTH Adam groups
  (homophily_learner.py:37-44; the conv parameters belong to both, homophily_agent.py:127-146) are averaged with one
  flat-bucket NCCL all-reduce between ``backward()`` and the clipping + two optimiser steps of lines 220-226.

Everything else -- unrolling the MAC and the target MAC over the episode, double-Q targets, the two TD losses, the similarity
loss, clipping order, optimiser order, target-network period, logged keys -- follows the reference step by step.
"""
from __future__ import annotations

import copy

import torch
from torch.optim import Adam

from .incentive import incentive_rewards

_NEG = -9999999.0                                             # homophily_learner.py:131-132


def _clone_mac(mac):
    """``copy.deepcopy(mac)`` (homophily_learner.py:47) that also works on a MAC which has already been rolled out: hidden states
    and cached inputs carrying autograd history cannot be deep-copied and are re-created by ``init_hidden`` anyway."""
    parked = {k: v for k, v in vars(mac).items() if torch.is_tensor(v) and v.grad_fn is not None}
    for k in parked:
        object.__setattr__(mac, k, None)
    try:
        return copy.deepcopy(mac)
    finally:
        for k, v in parked.items():
            object.__setattr__(mac, k, v)


class FlatBucket:
    """One contiguous fp32 buffer over the gradients of a parameter list: a single all-reduce per optimiser step."""

    def __init__(self, params, group=None):
        seen, self.params = set(), []
        for p in params:                                      # a parameter shared by both Adam groups is reduced once
            if id(p) not in seen and p.requires_grad:
                seen.add(id(p))
                self.params.append(p)
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def all_reduce_mean(self):
        import torch.distributed as dist
        world = dist.get_world_size(self.group)
        if world == 1:
            return
        dev = self.params[0].device
        if self.flat is None or self.flat.device != dev:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(list(self.flat.split([p.numel() for p in self.params])), [g.reshape(-1) for g in grads])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(world)
        for p, chunk in zip(self.params, self.flat.split([p.numel() for p in self.params])):
            if p.grad is None:
                p.grad = chunk.view_as(p).clone()
            else:
                p.grad.copy_(chunk.view_as(p))

    def broadcast_params(self, src=0):
        import torch.distributed as dist
        if dist.get_world_size(self.group) == 1:
            return
        for p in self.params:
            dist.broadcast(p.data, src=src, group=self.group)


def shard_episodes(batch, rank: int, world: int):
    """Contiguous episode shard of a sampled ``EpisodeBatch`` (its own slicing, episode_buffer.py:147-160)."""
    from .sharding import shard_range
    lo, hi = shard_range(batch.batch_size, rank, world)
    return batch[lo:hi]


class DeviceHomophilyLearner:
    def __init__(self, mac, scheme, logger, args, process_group=None):
        self.args, self.mac, self.logger = args, mac, logger
        self.device = args.device
        self.n_agents, self.n_actions = args.n_agents, args.n_actions
        n, dev = self.n_agents, self.device
        off_diag = 1 - torch.eye(n)
        self.inc_mask = off_diag.reshape(1, 1, n, n).to(dev)
        self.sim_mask_ik = off_diag.reshape(1, 1, n, n, 1).to(dev)      # i != k
        self.sim_mask_ij = off_diag.reshape(1, 1, n, 1, n).to(dev)      # i != j
        self.sim_mask_kj = off_diag.reshape(1, 1, 1, n, n).to(dev)      # k != j
        self.sim_horizon = args.sim_horizon
        self.params = list(mac.parameters())
        self.params_env = mac.parameters_env()
        self.params_inc = mac.parameters_inc()
        self.last_target_update_episode = 0
        self.optimiser_env = Adam(params=self.params_env, lr=args.lr_env)
        self.optimiser_inc = Adam(params=self.params_inc, lr=args.lr_inc)
        self.target_mac = _clone_mac(mac)
        self.log_stats_t = -self.args.learner_log_interval - 1
        self.group = process_group
        self.bucket = FlatBucket(list(self.params_inc) + list(self.params_env), group=process_group)
        self._dist = False
        try:
            import torch.distributed as dist
            self._dist = dist.is_available() and dist.is_initialized()
        except Exception:
            pass
        self._synced = False                                   # rank 0's parameters are broadcast once they live on the device

    # ------------------------------------------------------------------ pieces of cal_loss_and_step
    def _unroll(self, mac, batch, detach=False):
        """q_env [bs, t, n, A], q_inc [bs, t, n, n, 3] over the whole episode (homophily_learner.py:68-88)."""
        qe, qi = [], []
        mac.init_hidden(batch.batch_size)
        for t in range(batch.max_seq_length):
            e, i, _ = mac.forward(batch, t=t)
            qe.append(e.detach() if detach else e)
            qi.append(i.detach() if detach else i)
        return torch.stack(qe, dim=1), torch.stack(qi, dim=1)

    def _incentive(self, actions_inc, rewards, L):
        """homophily_learner.py:98-115 on the device: one kernel."""
        if rewards.is_cuda:
            r_env, r_inc, _ = incentive_rewards(actions_inc, rewards, float(self.args.incentive), float(self.args.incentive_cost),
                                                float(self.args.incentive_ratio), L, recip=True)
            return r_env, r_inc
        raise RuntimeError("DeviceHomophilyLearner needs CUDA tensors; there is no CPU fallback")

    @staticmethod
    def _receive_counts(actions_inc_all, n):
        off = (1 - torch.eye(n, device=actions_inc_all.device)).reshape(1, 1, n, n, 1)
        masked = actions_inc_all * off.long()
        pos = (masked == 1).sum(dim=(2, 4))
        neg = (masked == 2).sum(dim=(2, 4))
        give = (masked != 0).sum(dim=(3, 4))
        return pos, neg, n - 1 - pos - neg, give

    def _similarity(self, rewards, clean):
        """[bs, t-1, n, n, 1]: agents i, k were 'doing the same thing' over the last sim_horizon steps (lines 174-203)."""
        h = self.sim_horizon
        c_cum, r_cum = torch.cumsum(clean, dim=1), torch.cumsum(rewards, dim=1)
        c_h, r_h = c_cum.clone(), r_cum.clone()
        c_h[:, h:] -= c_cum[:, :-h]
        r_h[:, h:] -= r_cum[:, :-h]
        c_t, r_t = (c_h > 0).float(), (r_h > 0).float()
        cluster = 2 * r_t + c_t                                # a cluster is a distinct point of {0,1}^2 (see module docstring)
        idle = c_t + r_t
        both = (idle.unsqueeze(2) * idle.unsqueeze(3)).unsqueeze(-1)
        return (cluster.unsqueeze(2) == cluster.unsqueeze(3)).unsqueeze(-1).float() * both

    def losses(self, batch):
        """The three losses of one learner step and the quantities the reference logs (no optimiser step)."""
        a = self.args
        L, n = batch.max_seq_length, self.n_agents
        rewards = batch["reward"][:, :-1] / a.reward_scale
        actions = batch["actions"][:, :-1]
        actions_inc_all = batch["actions_inc"]
        actions_inc = actions_inc_all[:, :-1]
        clean = (batch["clean_num"][:, :-1] > 0).float()
        terminated = batch["terminated"][:, :-1].float()
        mask = batch["filled"][:, :-1].float()
        mask[:, 1:] = mask[:, 1:] * (1 - terminated[:, :-1])
        avail = batch["avail_actions"]

        q_env, q_inc = self._unroll(self.mac, batch)
        tq_env, tq_inc = self._unroll(self.target_mac, batch, detach=True)
        tq_env, tq_inc = tq_env[:, 1:].clone(), tq_inc[:, 1:].clone()

        r_env, r_inc = self._incentive(actions_inc, rewards, L)

        chosen_env = torch.gather(q_env[:, :-1], dim=-1, index=actions)
        others = bool(getattr(a, "consider_others_inc", False))
        if others:
            pos, neg, zero, _ = self._receive_counts(actions_inc_all, n)
            chosen_inc = (q_inc[:, :-1, :, :, 0] * zero[:, :-1].unsqueeze(2) + q_inc[:, :-1, :, :, 1] * pos[:, :-1].unsqueeze(2)
                          + q_inc[:, :-1, :, :, 2] * neg[:, :-1].unsqueeze(2)) / (n - 1)
        else:
            chosen_inc = torch.gather(q_inc[:, :-1], dim=-1, index=actions_inc).squeeze(-1)

        tq_env[avail[:, 1:] == 0] = _NEG
        if a.double_q:
            live_env = q_env.detach().clone()
            live_env[avail == 0] = _NEG
            best_env = live_env[:, 1:].max(dim=-1, keepdim=True)[1]
            best_inc = q_inc.detach()[:, 1:].max(dim=-1, keepdim=True)[1]
            t_env = torch.gather(tq_env, dim=-1, index=best_env)
            t_inc = torch.gather(tq_inc, dim=-1, index=best_inc).squeeze(-1)
        else:
            t_env = tq_env.max(dim=-1)[0]
            t_inc = tq_inc.max(dim=-1)[0].squeeze(-1)
        if others:
            t_other = (tq_inc[..., 0] * zero[:, 1:].unsqueeze(2) + tq_inc[..., 1] * pos[:, 1:].unsqueeze(2)
                       + tq_inc[..., 2] * neg[:, 1:].unsqueeze(2))
            t_next = torch.gather(tq_inc, dim=-1, index=actions_inc_all[:, 1:]).squeeze(-1)
            t_inc = (t_inc + t_other - t_next) / (n - 1)

        targets_env = r_env + a.gamma_env * (1 - terminated) * t_env.sum(dim=-1)
        targets_inc = r_inc + a.gamma_inc * (1 - terminated) * (t_inc * self.inc_mask).sum(dim=-1)
        td_env = chosen_env.sum(dim=-1) - targets_env.detach()
        td_inc = (chosen_inc * self.inc_mask).sum(dim=-1) - targets_inc.detach()
        m = mask.expand_as(td_env)
        loss_env = ((td_env * m) ** 2).sum() / m.sum()
        loss_inc = ((td_inc * m) ** 2).sum() / m.sum()

        sim = self._similarity(rewards, clean)
        probs = torch.softmax(q_inc, dim=-1)[:, :-1]                                    # [bs, t-1, i, j, a]
        by_k = actions_inc.unsqueeze(2).expand(-1, -1, n, -1, -1, -1)                   # [bs, t-1, (i), k, j, 1]
        p_ikj = torch.gather(probs.unsqueeze(3).expand(-1, -1, -1, n, -1, -1), dim=-1, index=by_k).squeeze(-1)
        sim_mask = torch.relu(sim.detach()) * self.sim_mask_ik * self.sim_mask_ij * self.sim_mask_kj
        loss_sim = (torch.clamp_min(-torch.log(p_ikj), a.sim_threshold) * sim_mask).sum() / (1 + sim_mask.sum())

        logs = {"loss_value_env": loss_env, "loss_value_inc": loss_inc, "loss_sim": loss_sim}
        with torch.no_grad():
            pos, neg, _, give = self._receive_counts(actions_inc_all, n)
            recv = (pos - neg)[:, :-1]
            logs["incentives_to_cleanup_per"] = (clean * recv).sum() / (clean.sum() + 1e-6)
            logs["incentives_to_harvest_per"] = (rewards * recv).sum() / (rewards.sum() + 1e-6)
            logs["value_give_mean"] = give[:, :-1].float().mean()
            logs["value_receive_mean"] = recv.float().mean()
            logs["q_env_taken_mean"] = chosen_env.squeeze(-1).mean()
            logs["q_inc_taken_mean"] = torch.gather(q_inc[:, :-1], dim=-1, index=actions_inc).squeeze(-1).mean()
        return loss_env, loss_inc, loss_sim, logs

    def _sync_start(self):
        """All ranks start from rank 0's parameters (the reference builds the learner on the CPU and moves it afterwards,
        run.py:132-135, so this cannot happen in __init__ with NCCL)."""
        if self._dist and not self._synced:
            self.bucket.broadcast_params(0)
            self.target_mac.load_state(self.mac)
        self._synced = True

    def cal_loss_and_step(self, batch):
        self._sync_start()
        loss_env, loss_inc, loss_sim, logs = self.losses(batch)
        self.optimiser_inc.zero_grad()
        self.optimiser_env.zero_grad()
        (loss_inc + loss_env + loss_sim * self.args.sim_loss_weight).backward()
        if self._dist:
            self.bucket.all_reduce_mean()                      # one flat bucket: both Adam groups, conv parameters once
        torch.nn.utils.clip_grad_norm_(self.params_inc, self.args.grad_norm_clip)       # order as lines 222-225
        torch.nn.utils.clip_grad_norm_(self.params_env, self.args.grad_norm_clip)
        self.optimiser_inc.step()
        self.optimiser_env.step()
        return logs

    # ------------------------------------------------------------------ reference surface
    def train(self, batch, t_env: int, episode_num: int):
        if self._dist:
            import torch.distributed as dist
            batch = shard_episodes(batch, dist.get_rank(self.group), dist.get_world_size(self.group))
        clean_num = batch["clean_num"][:, :-1]
        apple_den = batch["apple_den"][:, :-1]
        logs = self.cal_loss_and_step(batch)
        if (episode_num - self.last_target_update_episode) / self.args.target_update_interval >= 1.0:
            self._update_targets()
            self.last_target_update_episode = episode_num
        if t_env - self.log_stats_t >= self.args.learner_log_interval:
            self.logger.log_stat("clean_num_mean", clean_num.mean().item(), t_env)
            self.logger.log_stat("apple_den_mean", apple_den.mean().item(), t_env)
            for k, v in logs.items():
                self.logger.log_stat(k, v.item(), t_env)
            self.log_stats_t = t_env

    def _update_targets(self):
        self.target_mac.load_state(self.mac)
        self.logger.console_logger.info("Updated target network")

    def cuda(self):
        self.mac.cuda()
        self.target_mac.cuda()
        self._sync_start()

    def save_models(self, path):
        self.mac.save_models(path)
        torch.save(self.optimiser_env.state_dict(), "{}/opt_env.th".format(path))
        torch.save(self.optimiser_inc.state_dict(), "{}/opt_inc.th".format(path))

    def load_models(self, path):
        self.mac.load_models(path)
        self.target_mac.load_models(path)
        self.optimiser_env.load_state_dict(torch.load("{}/opt_env.th".format(path), map_location=lambda storage, loc: storage))
        self.optimiser_inc.load_state_dict(torch.load("{}/opt_inc.th".format(path), map_location=lambda storage, loc: storage))


REGISTRY = {"homophily_learner_b200": DeviceHomophilyLearner}
