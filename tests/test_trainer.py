"""ImpalaTrainer drop-in: ring layout / argument validation on CPU; on the GPU a short actor+learner run, the slot unpack,
the pipelined get_batch/learn loop against the synchronous learner, the versioned weight publish and the LSTM path."""
import math

import pytest
import torch

from scalerl_b200.algorithms.impala.impala_atari import ImpalaArguments, ImpalaTrainer, TrajectoryRing, slot_layout


def test_ring_schema_matches_create_buffers():
    T, A, nb = 5, 6, 3
    ring = TrajectoryRing(T, A, nb)
    want = dict(obs=((T + 1, 4, 84, 84), torch.uint8), reward=((T + 1,), torch.float32), done=((T + 1,), torch.bool),
                last_action=((T + 1,), torch.int64), action=((T + 1,), torch.int64), episode_return=((T + 1,), torch.float32),
                episode_step=((T + 1,), torch.int32), policy_logits=((T + 1, A), torch.float32), baseline=((T + 1,), torch.float32))
    assert set(ring.buffers) == set(want)                      # key schema of impala_atari.py:135-147
    for k, (shp, dt) in want.items():
        assert len(ring.buffers[k]) == nb
        assert tuple(ring.buffers[k][0].shape) == shp and ring.buffers[k][0].dtype == dt
        assert ring.buffers[k][0].is_shared()
    # slots do not alias: writing one slot/key leaves the others untouched
    ring.buffers['reward'][1][2] = 7.0
    ring.buffers['obs'][2][0, 0, 0, 0] = 9
    assert ring.buffers['reward'][0].sum() == 0 and ring.buffers['reward'][2].sum() == 0
    assert ring.buffers['obs'][1].sum() == 0 and ring.buffers['action'][2].sum() == 0
    lay, nbytes = slot_layout(T, A)
    assert nbytes % 64 == 0 and lay['obs'][0] == 0


def test_argument_validation_like_reference(tmp_path):
    a = ImpalaArguments(num_actors=4, num_buffers=4, batch_size=2, output_dir=str(tmp_path))
    with pytest.raises(ValueError):                            # impala_atari.py:74-75
        ImpalaTrainer(a)
    a = ImpalaArguments(num_actors=1, num_buffers=2, batch_size=4, output_dir=str(tmp_path))
    with pytest.raises(ValueError):                            # impala_atari.py:76-77
        ImpalaTrainer(a)
    a = ImpalaArguments(num_actors=2, batch_size=6, output_dir=str(tmp_path))
    t = ImpalaTrainer(a)
    assert a.num_buffers == 6                                  # default applied BEFORE buffers are created
    assert len(t.buffers['obs']) == 6


def _fill_slots(t, a, seed=0):
    g = torch.Generator().manual_seed(seed)
    for m in range(a.num_buffers):
        t.buffers['obs'][m].copy_(torch.randint(0, 256, t.buffers['obs'][m].shape, dtype=torch.uint8, generator=g))
        t.buffers['reward'][m].copy_(torch.randn(t.buffers['reward'][m].shape, generator=g))
        t.buffers['done'][m].copy_(torch.rand(t.buffers['done'][m].shape, generator=g) < 0.3)
        t.buffers['action'][m].copy_(torch.randint(0, a.num_actions, t.buffers['action'][m].shape, generator=g))
        t.buffers['policy_logits'][m].copy_(torch.randn(t.buffers['policy_logits'][m].shape, generator=g))
        t.buffers['episode_return'][m].copy_(torch.randn(t.buffers['episode_return'][m].shape, generator=g))
        for st in t.rnn_state_buffers[m]:
            st.copy_(torch.randn(st.shape, generator=g) * 0.1)


def _queues():
    ctx = torch.multiprocessing.get_context('fork')
    return ctx.SimpleQueue(), ctx.SimpleQueue()


@pytest.mark.gpu
def test_train_short_run_on_gpu(tmp_path):
    a = ImpalaArguments(num_actors=2, batch_size=4, rollout_length=5, total_steps=4 * 5 * 6, output_dir=str(tmp_path), num_actions=6)
    t = ImpalaTrainer(a)
    w0 = t.actor_model.state_dict()['fc.weight'].clone()
    out = t.train()
    assert out['steps'] >= a.total_steps
    for k in ImpalaTrainer.stat_keys:
        assert k in out
    assert math.isfinite(out['total_loss'])
    # weights were published into the shared actor model: after the final flush the actors hold exactly the learner's weights
    assert out['weights_version'] == 6
    for n, v in t.learner.state_dict().items():
        assert torch.equal(t.actor_model.state_dict()[n], v.cpu()), n
    assert not torch.equal(w0, t.actor_model.state_dict()['fc.weight'])
    ck = torch.load(tmp_path / a.project / 'model.tar', weights_only=False)
    assert set(ck) == {'model_state_dict', 'optimizer_state_dict', 'hparam'}   # impala_atari.py:506-511
    assert set(ck['model_state_dict']) == {'conv1.weight', 'conv1.bias', 'conv2.weight', 'conv2.bias', 'conv3.weight', 'conv3.bias',
                                           'fc.weight', 'fc.bias', 'policy.weight', 'policy.bias', 'baseline.weight', 'baseline.bias'}
    assert set(ck['optimizer_state_dict']) == {'state', 'param_groups'} and len(ck['optimizer_state_dict']['state']) == 12   # torch.optim layout


@pytest.mark.gpu
def test_get_batch_slot_unpack_matches_stack(tmp_path):
    """get_batch (one H2D copy per slot + srl_unpack_slots) == the reference's torch.stack(dim=1) of the slot tensors;
    it returns WITHOUT waiting for the GPU and the slots come back once their copy has finished"""
    a = ImpalaArguments(num_actors=1, batch_size=3, rollout_length=4, num_buffers=5, output_dir=str(tmp_path), num_actions=6)
    t = ImpalaTrainer(a)
    _fill_slots(t, a)
    free_q, full_q = _queues()
    order = [4, 0, 2]
    for m in order:
        full_q.put(m)
    batch, state = t.get_batch(free_q, full_q)
    torch.cuda.synchronize()
    for k in ('obs', 'reward', 'done', 'action', 'policy_logits', 'episode_return'):
        want = torch.stack([t.buffers[k][m] for m in order], dim=1)          # impala_atari.py:248-251
        assert torch.equal(batch[k].cpu(), want), k
    t.flush()
    assert sorted(free_q.get() for _ in range(3)) == sorted(order)            # slots released after their copy
    assert state == tuple()


@pytest.mark.gpu
@pytest.mark.parametrize('use_lstm', [False, True])
def test_pipelined_loop_equals_synchronous_learner(tmp_path, use_lstm):
    """get_batch -> learn (asynchronous copies, lagged stats, asynchronous versioned publish) over several steps produces
    the same weights as feeding the same batches to a synchronous B200ImpalaLearner; every published version lands in the
    shared actor parameters; initial LSTM states travel through rnn_state_buffers (impala_atari.py:108-120,252-253)"""
    from scalerl_b200.learner import B200ImpalaLearner
    from tests.helpers import rel_l2
    T, B, A, steps = 4, 3, 6, 5
    a = ImpalaArguments(num_actors=1, batch_size=B, rollout_length=T, num_buffers=2 * B, output_dir=str(tmp_path), num_actions=A,
                        use_lstm=use_lstm)
    t = ImpalaTrainer(a)
    sd0 = {k: v.clone() for k, v in t.actor_model.state_dict().items()}
    free_q, full_q = _queues()
    ref = B200ImpalaLearner(t.hparams(), init_state_dict=sd0, process_group=False)
    all_stats, ref_stats = [], []
    for k in range(steps):
        _fill_slots(t, a, seed=100 + k)
        order = [(k + i) % a.num_buffers for i in range(B)]
        for m in order:
            full_q.put(m)
        want = {key: torch.stack([t.buffers[key][m] for m in order], dim=1).cuda() for key in ('obs', 'reward', 'done', 'action', 'policy_logits', 'episode_return')}
        want_state = tuple(torch.cat([t.rnn_state_buffers[m][i] for m in order], dim=1).cuda() for i in range(2)) if use_lstm else ()
        batch, state = t.get_batch(free_q, full_q, t.buffers, t.rnn_state_buffers)
        assert len(state) == (2 if use_lstm else 0)
        all_stats.append(t.learn(t.actor_model, None, batch, state))
        ref_stats.append(ref.learn(want, want_state))
        t.flush()                                  # the test refills the slots on the host: wait for the copies first
        while not free_q.empty():
            free_q.get()
        assert int(t.weights_version[0]) == k + 1
    # stats lag one step behind (the first call returns its own)
    assert abs(all_stats[0]['total_loss'] - ref_stats[0]['total_loss']) <= 1e-4 * max(1.0, abs(ref_stats[0]['total_loss']))
    for k in range(2, steps):
        # two learners drift apart by fp32 summation order (RMSprop normalises near-zero gradients to +-lr): loose on purpose
        assert abs(all_stats[k]['total_loss'] - ref_stats[k - 1]['total_loss']) <= 1e-2 * max(1.0, abs(ref_stats[k - 1]['total_loss'])), k
        assert all_stats[k]['episode_returns'] == ref_stats[k - 1]['episode_returns']
    for n, v in ref.state_dict().items():
        assert rel_l2(t.learner.params[n].cpu(), v.cpu()) < 1e-2, n       # two learners drift by fp32 summation order; see above
        assert torch.equal(t.actor_model.state_dict()[n], t.learner.params[n].cpu()), n     # published == learner


@pytest.mark.gpu
def test_publish_skips_non_finite_weights(tmp_path):
    """a step whose loss is NaN/Inf never reaches the actors: the device-side snapshot keeps the last good weights"""
    a = ImpalaArguments(num_actors=1, batch_size=2, rollout_length=3, num_buffers=4, output_dir=str(tmp_path))
    t = ImpalaTrainer(a)
    t._ensure_learner()
    good = {k: v.clone() for k, v in t.actor_model.state_dict().items()}
    v1 = t.publish_weights(t.actor_model)
    assert v1 == 1 and int(t.weights_version[0]) == 1
    for n in good:
        assert torch.equal(t.actor_model.state_dict()[n], good[n])
    t.learner.flat_params.fill_(float('nan'))
    t.learner._losses.fill_(float('nan'))
    t.publish_weights(t.actor_model)
    assert int(t.weights_version[0]) == 2
    for n in good:
        assert torch.equal(t.actor_model.state_dict()[n], good[n]), n


@pytest.mark.gpu
@pytest.mark.parametrize('optimizer', ['rmsprop', 'adam'])
def test_checkpoint_round_trip_restores_optimizer_step(tmp_path, optimizer):
    """save_checkpoint / load_checkpoint (algorithms/base.py:102-116, impala_atari.py:496-515): weights, moments AND the step
    count (Adam's bias correction) -- a resumed learner takes the same next step as the one that kept running"""
    from oracle import impala_oracle as O
    from scalerl_b200.learner import B200ImpalaLearner, ImpalaHParams
    T, B, A = 3, 4, 6
    hp = ImpalaHParams(rollout_length=T, batch_size=B, num_actions=A, optimizer=optimizer, learning_rate=1e-3)
    L = B200ImpalaLearner(hp, init_state_dict=O.init_params(A, seed=1), process_group=False)
    batches = [{k: v.cuda() for k, v in O.synthetic_batch(T, B, A, seed=s).items()} for s in range(4)]
    for b in batches[:3]:
        L.learn(b)
    assert L.global_opt_step == 3 and L.device_opt_step() == 3
    path = str(tmp_path / 'ck.tar')
    L.save_checkpoint(path)
    ck = torch.load(path, weights_only=False)
    assert all(float(s['step']) == 3.0 for s in ck['optimizer_state_dict']['state'].values())
    R = B200ImpalaLearner(hp, init_state_dict=O.init_params(A, seed=2), process_group=False)
    R.load_checkpoint(path)
    assert R.global_opt_step == 3 and R.device_opt_step() == 3
    assert torch.equal(R.flat_params, L.flat_params) and torch.equal(R.opt_state0, L.opt_state0)
    L.learn(batches[3]); R.learn(batches[3])
    assert L.device_opt_step() == 4 and R.device_opt_step() == 4
    rel = float((R.flat_params - L.flat_params).norm() / L.flat_params.norm())
    assert rel < 1e-6, rel
    # the torch optimizer of the reference loads the same file
    net = torch.nn.ParameterList([torch.nn.Parameter(v.clone()) for v in ck['model_state_dict'].values()])
    opt = torch.optim.RMSprop(net, lr=1e-3, alpha=0.99, eps=1e-5) if optimizer == 'rmsprop' else torch.optim.Adam(net, lr=1e-3)
    opt.load_state_dict(ck['optimizer_state_dict'])
