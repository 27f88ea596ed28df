"""ImpalaTrainer drop-in: ring layout / argument validation on CPU, a short actor+learner run on the GPU."""
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
    with pytest.raises(NotImplementedError):
        ImpalaTrainer(ImpalaArguments(use_lstm=True, output_dir=str(tmp_path)))


@pytest.mark.gpu
def test_train_short_run_on_gpu(tmp_path):
    a = ImpalaArguments(num_actors=2, batch_size=4, rollout_length=5, total_steps=4 * 5 * 6, output_dir=str(tmp_path), num_actions=6)
    t = ImpalaTrainer(a)
    w0 = t.actor_model.reference_state_dict()['fc.weight'].clone()
    out = t.train()
    assert out['steps'] >= a.total_steps
    for k in ImpalaTrainer.stat_keys:
        assert k in out
    assert torch.isfinite(torch.tensor(out['total_loss']))
    w1 = t.actor_model.reference_state_dict()['fc.weight']
    assert not torch.equal(w0, w1)                              # weights were published to the shared actor model
    ck = torch.load(tmp_path / a.project / 'model.tar', weights_only=False)
    assert set(ck) == {'model_state_dict', 'optimizer_state_dict', 'hparam'}   # impala_atari.py:506-511
    assert set(ck['model_state_dict']) == {'conv1.weight', 'conv1.bias', 'conv2.weight', 'conv2.bias', 'conv3.weight', 'conv3.bias',
                                           'fc.weight', 'fc.bias', 'policy.weight', 'policy.bias', 'baseline.weight', 'baseline.bias'}


@pytest.mark.gpu
def test_get_batch_slot_unpack_matches_stack(tmp_path):
    """get_batch (one H2D copy per slot + srl_unpack_slots) == the reference's torch.stack(dim=1) of the slot tensors"""
    a = ImpalaArguments(num_actors=1, batch_size=3, rollout_length=4, num_buffers=5, output_dir=str(tmp_path), num_actions=6)
    t = ImpalaTrainer(a)
    g = torch.Generator().manual_seed(0)
    for m in range(a.num_buffers):
        t.buffers['obs'][m].copy_(torch.randint(0, 256, t.buffers['obs'][m].shape, dtype=torch.uint8, generator=g))
        t.buffers['reward'][m].copy_(torch.randn(t.buffers['reward'][m].shape, generator=g))
        t.buffers['done'][m].copy_(torch.rand(t.buffers['done'][m].shape, generator=g) < 0.3)
        t.buffers['action'][m].copy_(torch.randint(0, 6, t.buffers['action'][m].shape, generator=g))
        t.buffers['policy_logits'][m].copy_(torch.randn(t.buffers['policy_logits'][m].shape, generator=g))
        t.buffers['episode_return'][m].copy_(torch.randn(t.buffers['episode_return'][m].shape, generator=g))
    ctx = torch.multiprocessing.get_context('fork')
    free_q, full_q = ctx.SimpleQueue(), ctx.SimpleQueue()
    order = [4, 0, 2]
    for m in order:
        full_q.put(m)
    batch, state = t.get_batch(free_q, full_q)
    torch.cuda.synchronize()
    for k in ('obs', 'reward', 'done', 'action', 'policy_logits', 'episode_return'):
        want = torch.stack([t.buffers[k][m] for m in order], dim=1)          # impala_atari.py:248-251
        assert torch.equal(batch[k].cpu(), want), k
    assert sorted(free_q.get() for _ in range(3)) == sorted(order)            # slots released after their copy
    assert state == tuple()
