"""B200ActorModel -- batched GPU actor inference behind the reference's actor calling convention.

The reference's actors each run ``AtariNet`` on the CPU for ONE environment per call
(/root/reference scalerl/algorithms/impala/impala_atari.py:177-197).  This module evaluates the same network for N environments in
one call on the sm_100a forward kernels (``srl_learner_forward``: space-to-depth, the three tcgen05 convs, fc, heads) and samples the
actions on the device (``srl_sample_actions``) -- SEED-style central inference, but behind the SAME interface:

    agent_output, agent_state = actor_model(env_output, agent_state)        # env_output tensors are [1, N, ...]
    actor_model.initial_hidden_state(batch_size)   state_dict()   load_state_dict()   train() / eval()

so ``ImpalaTrainer.get_action``-style loops (and ``get_action_batched`` for N environments per actor process) use it unchanged.
Weights: its own flat fp32 buffer in the learner's layout; ``refresh(shared_flat, version)`` pulls a newer published version from
the shared-memory actor parameters with one H2D copy (the versioned publish of ImpalaTrainer.publish_weights), ``sync_from(learner)``
copies device-to-device when actor and learner share a process.  The LSTM core is not covered (rows are evaluated independently)."""
from collections import OrderedDict
from typing import Dict, Optional

import torch

from ... import _lib
from ...learner import B200ImpalaLearner, ImpalaHParams


class B200ActorModel:
    def __init__(self, num_envs: int, num_actions: int = 6, device=None, init_state_dict: Optional[Dict[str, torch.Tensor]] = None, seed: int = 0):
        # a forward-only context: T = 1 gives room for the single row an actor step evaluates (rows <= T + 1)
        self._ctx = B200ImpalaLearner(ImpalaHParams(rollout_length=1, batch_size=num_envs, num_actions=num_actions), device=device,
                                      process_group=False, init_state_dict=init_state_dict, seed=seed, use_graph=False)
        self.device = self._ctx.device
        self.num_envs, self.num_actions = num_envs, num_actions
        self.training = True
        self.weights_version = 0
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(seed)
        B = num_envs
        self._pin = {'obs': torch.empty(1, B, 4, 84, 84, dtype=torch.uint8).pin_memory(), 'reward': torch.empty(1, B).pin_memory(),
                     'action': torch.empty(1, B, dtype=torch.int64).pin_memory()}
        self._dev = {k: torch.empty_like(v, device=self.device) for k, v in self._pin.items()}
        self._out_host = {'policy_logits': torch.empty(1, B, num_actions).pin_memory(), 'baseline': torch.empty(1, B).pin_memory(),
                          'action': torch.empty(1, B, dtype=torch.int64).pin_memory()}
        self._flat_pin = None

    # ---- module-like surface ---------------------------------------------------------------------------------------------
    def train(self, mode: bool = True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    def share_memory(self):
        return self

    def initial_hidden_state(self, batch_size: int):
        return tuple()

    def state_dict(self):
        return OrderedDict((k, v.cpu()) for k, v in self._ctx.state_dict().items())

    def load_state_dict(self, sd, strict=True):
        self._ctx.load_state_dict(sd)

    # ---- weights -----------------------------------------------------------------------------------------------------------
    def sync_from(self, learner: B200ImpalaLearner, version: Optional[int] = None) -> None:
        """same process: device-to-device copy of the learner's flat parameters (same layout by construction)"""
        if learner.numel != self._ctx.numel:
            raise ValueError('learner and actor were built for different num_actions')
        self._ctx.flat_params.copy_(learner.flat_params, non_blocking=True)
        self.weights_version = self.weights_version + 1 if version is None else int(version)

    def refresh(self, shared_flat: torch.Tensor, version: int) -> bool:
        """pull a newer published version from the shared-memory flat actor parameters (ActorNet.flat_params): one H2D copy"""
        if int(version) == self.weights_version:
            return False
        if shared_flat.numel() != self._ctx.numel:
            raise ValueError('shared parameter buffer has another layout')
        if self._flat_pin is None:
            self._flat_pin = torch.empty(self._ctx.numel).pin_memory()
        self._flat_pin.copy_(shared_flat)                 # snapshot (the learner may be writing the next version)
        self._ctx.flat_params.copy_(self._flat_pin, non_blocking=True)
        self.weights_version = int(version)
        return True

    # ---- inference -----------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_device(self, obs: torch.Tensor, reward: torch.Tensor, action: torch.Tensor):
        """device tensors [1, N, ...] -> (policy_logits [1,N,A], baseline [1,N], action [1,N]) on the device, nothing synchronised"""
        out = self._ctx.forward({'obs': obs, 'reward': reward, 'action': action})
        lg = out['policy_logits']
        N = lg.shape[0] * lg.shape[1]
        act = torch.empty(lg.shape[0], lg.shape[1], dtype=torch.int64, device=self.device)
        u = torch.rand(N, device=self.device, generator=self._gen) if self.training else None
        _lib.check(_lib.lib().srl_sample_actions(lg.data_ptr(), u.data_ptr() if u is not None else None, N, self.num_actions, act.data_ptr(),
                                                 torch.cuda.current_stream(self.device).cuda_stream), 'srl_sample_actions')
        return lg, out['baseline'], act

    @torch.no_grad()
    def __call__(self, env_output: Dict[str, torch.Tensor], agent_state=()):
        """the reference's ``actor_model(env_output, agent_state)`` for N environments: host tensors in, host tensors out"""
        for k in ('obs', 'reward', 'action'):
            t = env_output[k]
            if t.is_cuda:
                self._dev[k].copy_(t, non_blocking=True)
            else:
                self._pin[k].copy_(t)
                self._dev[k].copy_(self._pin[k], non_blocking=True)
        lg, bs, act = self.forward_device(self._dev['obs'], self._dev['reward'], self._dev['action'])
        self._out_host['policy_logits'].copy_(lg, non_blocking=True)
        self._out_host['baseline'].copy_(bs, non_blocking=True)
        self._out_host['action'].copy_(act, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return {k: v.clone() for k, v in self._out_host.items()}, tuple()

    def close(self):
        self._ctx.close()
