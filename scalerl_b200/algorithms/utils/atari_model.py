"""CPU actor-side stand-ins used when ScaleRL's own AtariNet / gym env are not available.

``ActorNet`` evaluates the AtariNet architecture (conv 8s4 -> 4s2 -> 3s1 -> fc512 -> [h, clipped reward, one-hot
last action] -> [2-layer LSTM] -> policy/baseline heads; reference: scalerl/algorithms/utils/atari_model.py:30-59,
93-134) for the actor's one-step calls on the CPU and samples an action.  It follows the reference model's CALLING
CONVENTION -- ``actor_model(env_output, agent_state) -> (dict(policy_logits, baseline, action), agent_state)``,
``initial_hidden_state(batch_size)``, ``state_dict()`` / ``load_state_dict()`` with the reference's parameter names and
layouts -- so ImpalaTrainer treats it and the reference's own ``AtariNet`` identically (either can be passed as
``actor_model_fn``); its parameters live in shared memory and are overwritten by the learner's weight publish.
``SyntheticAtariEnv`` emits the TorchEnvWrapper record schema (scalerl/envs/torch_envwrapper.py:43-50,77-84)
with random frames; it exists so the actor/ring/learner plumbing can be exercised without gymnasium/ale_py.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from ...learner import param_shapes, reference_param_order


class ActorNet(torch.nn.Module):
    def __init__(self, obs_shape=(4, 84, 84), num_actions=6, use_lstm=False, seed=0):
        super().__init__()
        self.observation_shape = tuple(obs_shape)
        self.num_actions = int(num_actions)
        self.use_lstm = bool(use_lstm)
        self.core_size = 513 + self.num_actions
        self.names = reference_param_order(self.use_lstm)
        g = torch.Generator().manual_seed(seed)
        fan = 1
        shapes = param_shapes(num_actions, self.use_lstm)
        # All parameters are views of ONE flat fp32 buffer laid out like the learner's flat parameter buffer
        # (srl_param_layout_ex): the weight publish is then a single D2H copy into shared memory (impala_atari.py:348).
        from ... import _lib
        from ...learner import PARAM_NAMES, LSTM_PARAM_NAMES
        total, off, cnt = _lib.param_layout(self.num_actions, self.use_lstm)
        lay_names = PARAM_NAMES + (LSTM_PARAM_NAMES if self.use_lstm else ())
        self.flat_layout = {n: (off[i], cnt[i]) for i, n in enumerate(lay_names)}
        self.flat_params = torch.zeros(total)
        for n in self.names:
            shp = shapes[n]
            if n.startswith('rnn_layer.'):
                fan = self.core_size                       # nn.LSTM: U(+-1/sqrt(hidden_size)) for every tensor
            elif n.endswith('.weight'):
                fan = 1
                for d in shp[1:]:
                    fan *= d
            bound = 1.0 / fan ** 0.5
            o, c = self.flat_layout[n]
            view = self.flat_params[o:o + c].view(shp)
            view.copy_((torch.rand(shp, generator=g) * 2 - 1) * bound)
            self.register_parameter(n.replace('.', '_'), torch.nn.Parameter(view, requires_grad=False))

    def share_memory(self):
        """moves the ONE underlying storage to shared memory; every parameter stays a view of it"""
        self.flat_params.share_memory_()
        for n in self.names:
            o, c = self.flat_layout[n]
            self._p(n).data = self.flat_params[o:o + c].view(self._p(n).shape)
        return self

    def _p(self, n):
        return getattr(self, n.replace('.', '_'))

    # ---- the reference model's parameter interface (names of atari_model.py:30-59) -----------------------------------
    def state_dict(self, *args, **kwargs):
        return OrderedDict((n, self._p(n).detach()) for n in self.names)

    def load_state_dict(self, sd, strict=True):
        missing = [n for n in self.names if n not in sd]
        extra = [k for k in sd if k not in self.names]
        if strict and (missing or extra):
            raise RuntimeError(f'ActorNet.load_state_dict: missing keys {missing}, unexpected keys {extra}')
        with torch.no_grad():
            for n in self.names:
                if n in sd:
                    self._p(n).copy_(sd[n])

    def initial_hidden_state(self, batch_size: int):
        """atari_model.py:61-75: () without LSTM, else (h0, c0) zeros [2, batch, 513 + A]"""
        if not self.use_lstm:
            return tuple()
        return tuple(torch.zeros(2, batch_size, self.core_size) for _ in range(2))

    def _lstm_step(self, x, state):
        """one time step of the 2-layer nn.LSTM (gate order i, f, g, o); state = (h [2,B,H], c [2,B,H])"""
        h_in, c_in = state
        hs, cs = [], []
        for layer in (0, 1):
            gates = (F.linear(x, self._p(f'rnn_layer.weight_ih_l{layer}'), self._p(f'rnn_layer.bias_ih_l{layer}'))
                     + F.linear(h_in[layer], self._p(f'rnn_layer.weight_hh_l{layer}'), self._p(f'rnn_layer.bias_hh_l{layer}')))
            i, f, g, o = gates.chunk(4, dim=-1)
            c = torch.sigmoid(f) * c_in[layer] + torch.sigmoid(i) * torch.tanh(g)
            x = torch.sigmoid(o) * torch.tanh(c)
            hs.append(x)
            cs.append(c)
        return x, (torch.stack(hs), torch.stack(cs))

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
        if self.use_lstm:
            core = core.view(T, B, -1)
            notdone = (~inputs['done']).float()
            outs = []
            for t in range(T):
                nd = notdone[t].view(1, B, 1)
                rnn_state = tuple(nd * s for s in rnn_state)        # state reset at episode ends (atari_model.py:114-116)
                y, rnn_state = self._lstm_step(core[t], rnn_state)
                outs.append(y)
            core = torch.cat(outs, 0)
        else:
            rnn_state = tuple()
        logits = F.linear(core, self._p('policy.weight'), self._p('policy.bias'))
        baseline = F.linear(core, self._p('baseline.weight'), self._p('baseline.bias'))
        if self.training:
            action = torch.multinomial(F.softmax(logits, dim=1), num_samples=1)
        else:
            action = torch.argmax(logits, dim=1)
        return dict(policy_logits=logits.view(T, B, -1), baseline=baseline.view(T, B), action=action.view(T, B)), rnn_state


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
