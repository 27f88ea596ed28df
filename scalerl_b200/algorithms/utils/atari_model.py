"""CPU actor-side stand-ins used when ScaleRL's own AtariNet / gym env are not available.

``ActorNet`` evaluates the AtariNet architecture (conv 8s4 -> 4s2 -> 3s1 -> fc512 -> [h, clipped reward, one-hot
last action] -> policy/baseline heads; reference: scalerl/algorithms/utils/atari_model.py:30-59,93-134) for ONE
environment step on the CPU and samples an action; its parameters live in shared memory and are overwritten by the
learner's weight publish.  Parameter names/layouts are the reference's state_dict (so checkpoints interchange).
``SyntheticAtariEnv`` emits the TorchEnvWrapper record schema (scalerl/envs/torch_envwrapper.py:43-50,77-84)
with random frames; it exists so the actor/ring/learner plumbing can be exercised without gymnasium/ale_py.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from ...learner import PARAM_NAMES, param_shapes


class ActorNet(torch.nn.Module):
    def __init__(self, obs_shape=(4, 84, 84), num_actions=6, seed=0):
        super().__init__()
        self.num_actions = num_actions
        g = torch.Generator().manual_seed(seed)
        fan = 1
        for n, shp in param_shapes(num_actions).items():
            if n.endswith('.weight'):
                fan = 1
                for d in shp[1:]:
                    fan *= d
            bound = 1.0 / fan ** 0.5
            self.register_parameter(n.replace('.', '_'), torch.nn.Parameter((torch.rand(shp, generator=g) * 2 - 1) * bound,
                                                                           requires_grad=False))

    def _p(self, n):
        return getattr(self, n.replace('.', '_'))

    def reference_state_dict(self):
        return OrderedDict((n, self._p(n).detach().clone()) for n in PARAM_NAMES)

    def load_reference_state_dict(self, sd):
        with torch.no_grad():
            for n in PARAM_NAMES:
                self._p(n).copy_(sd[n])

    @torch.no_grad()
    def forward(self, inputs, rnn_state=()):
        x = inputs['obs']
        T, B = x.shape[:2]
        x = x.reshape(T * B, *x.shape[2:]).float() / 255.0
        x = F.relu(F.conv2d(x, self._p('conv1.weight'), self._p('conv1.bias'), stride=4))
        x = F.relu(F.conv2d(x, self._p('conv2.weight'), self._p('conv2.bias'), stride=2))
        x = F.relu(F.conv2d(x, self._p('conv3.weight'), self._p('conv3.bias'), stride=1))
        x = F.relu(F.linear(x.reshape(T * B, -1), self._p('fc.weight'), self._p('fc.bias')))
        one_hot = F.one_hot(inputs['action'].reshape(T * B), self.num_actions).float()
        core = torch.cat([x, torch.clamp(inputs['reward'], -1, 1).reshape(T * B, 1), one_hot], dim=-1)
        logits = F.linear(core, self._p('policy.weight'), self._p('policy.bias'))
        baseline = F.linear(core, self._p('baseline.weight'), self._p('baseline.bias'))
        action = torch.multinomial(F.softmax(logits, dim=1), num_samples=1)
        return dict(policy_logits=logits.view(T, B, -1), baseline=baseline.view(T, B), action=action.view(T, B))


class SyntheticAtariEnv:
    """random 84x84x4 uint8 frames, reward in {-1,0,1}, episodes of ~200 steps; record schema of TorchEnvWrapper"""

    def __init__(self, obs_shape=(4, 84, 84), num_actions=6, seed=0, episode_len=200):
        self.obs_shape, self.num_actions, self.episode_len = obs_shape, num_actions, episode_len
        self.g = torch.Generator().manual_seed(seed + 12345 + torch.initial_seed() % 1000)
        self.episode_return = torch.zeros(1, 1)
        self.episode_step = torch.zeros(1, 1, dtype=torch.int32)

    def _frame(self):
        return torch.randint(0, 256, (1, 1, *self.obs_shape), dtype=torch.uint8, generator=self.g)

    def reset(self):
        self.episode_return.zero_()
        self.episode_step.zero_()
        return dict(obs=self._frame(), reward=torch.zeros(1, 1), done=torch.ones(1, 1, dtype=torch.bool),
                    episode_return=self.episode_return.clone(), episode_step=self.episode_step.clone(),
                    action=torch.zeros(1, 1, dtype=torch.int64))

    def step(self, action):
        r = float(torch.randint(-1, 2, (1,), generator=self.g))
        self.episode_step += 1
        self.episode_return += r
        done = bool(self.episode_step.item() >= self.episode_len)
        out = dict(obs=self._frame(), reward=torch.full((1, 1), r), done=torch.tensor([[done]]),
                   episode_return=self.episode_return.clone(), episode_step=self.episode_step.clone(),
                   action=action.view(1, 1).to(torch.int64))
        if done:
            self.episode_return.zero_()
            self.episode_step.zero_()
        return out
