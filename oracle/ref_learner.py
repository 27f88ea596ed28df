"""The reference's CPU learner step, run on the reference's OWN modules (oracle/_ref, built by oracle/make_ref.py):
``AtariNet`` (atari_model.py), ``vtrace.from_logits`` and ``loss_fn.compute_*`` driven by the statements of
``ImpalaTrainer.learn`` (/root/reference scalerl/algorithms/impala/impala_atari.py:288-346) and the optimizer of
``setup_optimizer`` (:99-105) / ``clip_grad_norm_`` (:344-345).  TEST / BENCH INFRASTRUCTURE: used by ``bench.py --impl reference``,
the ``cpu_baseline`` leg and the tests; the product package never imports it."""
import importlib.util
import os

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ('vtrace.py', 'loss_fn.py', 'atari_model.py'))


def _load(name):
    spec = importlib.util.spec_from_file_location(f'srl_ref_{name}', os.path.join(REF_DIR, f'{name}.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ReferenceLearner:
    """learner_model + optimizer + learn(), hyper-parameters as ImpalaTrainer reads them (defaults: SURVEY.md §8d)"""

    def __init__(self, num_actions=6, use_lstm=False, state_dict=None, discounting=0.99, baseline_cost=0.5, entropy_cost=0.0006,
                 reward_clipping='abs_one', max_grad_norm=40.0, learning_rate=1e-4, alpha=0.99, epsilon=1e-5, momentum=0.0, seed=0):
        if not available():
            raise RuntimeError('oracle/_ref is missing: run python oracle/make_ref.py in the build container')
        self.vtrace, self.loss_fn, am = _load('vtrace'), _load('loss_fn'), _load('atari_model')
        torch.manual_seed(seed)
        self.model = am.AtariNet((4, 84, 84), num_actions, use_lstm=use_lstm)
        if state_dict is not None:
            self.model.load_state_dict(state_dict)
        self.optimizer = torch.optim.RMSprop(self.model.parameters(), lr=learning_rate, momentum=momentum, eps=epsilon, alpha=alpha)   # :99-105
        self.hp = dict(discounting=discounting, baseline_cost=baseline_cost, entropy_cost=entropy_cost, reward_clipping=reward_clipping,
                       max_grad_norm=max_grad_norm)

    def learn(self, batch, initial_rnn_state=()):
        """impala_atari.py:288-346 (without the lock and the actor weight copy)"""
        h = self.hp
        learner_outputs, unused_state = self.model(batch, initial_rnn_state)
        bootstrap_value = learner_outputs['baseline'][-1]
        batch = {key: tensor[1:] for key, tensor in batch.items()}
        learner_outputs = {key: tensor[:-1] for key, tensor in learner_outputs.items()}
        rewards = batch['reward']
        clipped_rewards = torch.clamp(rewards, -1, 1) if h['reward_clipping'] == 'abs_one' else rewards
        discounts = (~batch['done']).float() * h['discounting']
        vtrace_returns = self.vtrace.from_logits(behavior_policy_logits=batch['policy_logits'], target_policy_logits=learner_outputs['policy_logits'],
                                                 actions=batch['action'], discounts=discounts, rewards=clipped_rewards,
                                                 values=learner_outputs['baseline'], bootstrap_value=bootstrap_value)
        pg_loss = self.loss_fn.compute_policy_gradient_loss(learner_outputs['policy_logits'], batch['action'], vtrace_returns.pg_advantages)
        baseline_loss = h['baseline_cost'] * self.loss_fn.compute_baseline_loss(vtrace_returns.vs - learner_outputs['baseline'])
        entropy_loss = h['entropy_cost'] * self.loss_fn.compute_entropy_loss(learner_outputs['policy_logits'])
        total_loss = pg_loss + baseline_loss + entropy_loss
        episode_returns = batch['episode_return'][batch['done']]
        stats = {'episode_returns': tuple(episode_returns.cpu().numpy()), 'mean_episode_return': torch.mean(episode_returns).item(),
                 'total_loss': total_loss.item(), 'pg_loss': pg_loss.item(), 'baseline_loss': baseline_loss.item(),
                 'entropy_loss': entropy_loss.item()}
        self.optimizer.zero_grad()
        total_loss.backward()
        stats['grad_norm'] = float(nn.utils.clip_grad_norm_(self.model.parameters(), h['max_grad_norm']))
        self.optimizer.step()
        return stats
