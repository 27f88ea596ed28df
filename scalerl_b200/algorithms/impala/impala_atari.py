"""ImpalaTrainer -- drop-in for scalerl/algorithms/impala/impala_atari.py with the learner on a B200.

Same public surface as the reference class (``ImpalaTrainer(args)``, ``create_buffers``, ``get_action``,
``get_batch``, ``learn``, ``learn_process``, ``train``, ``save_checkpoint``; impala_atari.py:40-515), same
trajectory key schema (impala_atari.py:122-151), same stats keys (:333-340) and checkpoint keys (:506-511).
What changes behind it:

  * buffers live in ONE shared-memory block per slot (obs + the small fields), registered as pinned host
    memory by the learner process, so ``get_batch`` issues asynchronous H2D copies on a copy stream instead of
    ``torch.stack`` + a pageable ``.to(device)`` (impala_atari.py:248-265); a slot returns to ``free_queue`` only
    after its copy event fired (ownership rule);
  * ``learn`` runs the sm_100a kernels (B200ImpalaLearner) and publishes the new weights into the shared CPU
    actor parameters (impala_atari.py:348);
  * CUDA is first touched inside the learner process (the reference forks after building models in the parent,
    SURVEY.md §7 hard part 8); ``global_step`` is a shared counter (the reference's plain int never reaches the
    parent).

Actors stay ordinary Python processes running a CPU policy.  Inside ScaleRL they use the reference's own
``AtariNet`` + ``TorchEnvWrapper``; ``env_fn`` / ``actor_model_fn`` default to the self-contained stand-ins of
``scalerl_b200.algorithms.utils`` because gymnasium / ale_py are not installed in this image.
"""
from __future__ import annotations

import math
import os
import time
import timeit
import traceback
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
from torch import multiprocessing as mp

from ...learner import B200ImpalaLearner, ImpalaHParams, PARAM_NAMES
from ..utils.atari_model import ActorNet, SyntheticAtariEnv


@dataclass
class ImpalaArguments:
    """RLArguments fields ImpalaTrainer reads (rl_args.py:71-159) plus the ones it reads but RLArguments never
    defined (impala_atari.py:56,72-77,303-308,325-327,375,412,502) with upstream torchbeast defaults."""
    env_id: str = 'PongNoFrameskip-v4'
    project: str = 'impala'
    algo_name: str = 'impala_b200'
    output_dir: str = './work_dir'
    use_cuda: bool = True
    num_actors: int = 4
    num_learners: int = 1
    num_buffers: Optional[int] = None
    batch_size: int = 8
    rollout_length: int = 20
    total_steps: int = 100000
    use_lstm: bool = False
    reward_clipping: str = 'abs_one'
    discounting: float = 0.99
    baseline_cost: float = 0.5
    entropy_cost: float = 0.0006
    max_grad_norm: float = 40.0
    learning_rate: float = 1e-4
    alpha: float = 0.99
    momentum: float = 0.0
    epsilon: float = 1e-5
    optimizer: str = 'rmsprop'
    disable_checkpoint: bool = False
    num_actions: int = 6
    obs_shape: Tuple[int, int, int] = (4, 84, 84)
    seed: int = 0


def slot_layout(T: int, A: int, obs_shape=(4, 84, 84)):
    """byte layout of one trajectory slot: every key of create_buffers (impala_atari.py:135-147), 64-byte aligned"""
    n = T + 1
    obs_elems = 1
    for d in obs_shape:
        obs_elems *= d
    specs = [('obs', (n, *obs_shape), torch.uint8), ('reward', (n,), torch.float32), ('done', (n,), torch.bool),
             ('last_action', (n,), torch.int64), ('action', (n,), torch.int64), ('episode_return', (n,), torch.float32),
             ('episode_step', (n,), torch.int32), ('policy_logits', (n, A), torch.float32), ('baseline', (n,), torch.float32)]
    off, out = 0, {}
    for k, shp, dt in specs:
        nbytes = torch.empty(0, dtype=dt).element_size()
        for d in shp:
            nbytes *= d
        out[k] = (off, shp, dt, nbytes)
        off = (off + nbytes + 63) & ~63
    return out, off


class TrajectoryRing:
    """num_buffers slots in one shared-memory uint8 block; ``buffers[key][m]`` are typed views (the reference's
    ``buffers[key][index][t, ...] = value`` writes work unchanged).  ``pin()`` registers the block with CUDA."""

    def __init__(self, T: int, A: int, num_buffers: int, obs_shape=(4, 84, 84)):
        self.T, self.A, self.num_buffers = T, A, num_buffers
        self.layout, self.slot_bytes = slot_layout(T, A, obs_shape)
        self.block = torch.zeros(num_buffers * self.slot_bytes, dtype=torch.uint8).share_memory_()
        self.buffers: Dict[str, List[torch.Tensor]] = {k: [] for k in self.layout}
        for m in range(num_buffers):
            base = m * self.slot_bytes
            for k, (off, shp, dt, nbytes) in self.layout.items():
                self.buffers[k].append(self.block[base + off: base + off + nbytes].view(dt).view(shp))
        self._pinned = False

    def pin(self):
        if not self._pinned:
            rc = torch.cuda.cudart().cudaHostRegister(self.block.data_ptr(), self.block.numel(), 0)
            if int(rc) != 0:
                raise RuntimeError(f'cudaHostRegister failed: {rc}')
            self._pinned = True

    def unpin(self):
        if self._pinned:
            torch.cuda.cudart().cudaHostUnregister(self.block.data_ptr())
            self._pinned = False


class ImpalaTrainer:
    stat_keys = ['total_loss', 'mean_episode_return', 'pg_loss', 'baseline_loss', 'entropy_loss']

    def __init__(self, args: ImpalaArguments, env_fn: Optional[Callable[[], Any]] = None,
                 actor_model_fn: Optional[Callable[[], torch.nn.Module]] = None) -> None:
        self.args = args
        if args.use_lstm:
            raise NotImplementedError('use_lstm=True is the "next" row of SURVEY.md §8f; the B200 learner is the non-LSTM core')
        if args.num_buffers is None:                                   # impala_atari.py:72-73, applied BEFORE create_buffers
            args.num_buffers = max(2 * args.num_actors, args.batch_size)
        if args.num_actors >= args.num_buffers:                        # :74-75
            raise ValueError('num_buffers should be larger than num_actors')
        if args.num_buffers < args.batch_size:                         # :76-77
            raise ValueError('num_buffers should be larger than batch_size')
        self.env_fn = env_fn or (lambda: SyntheticAtariEnv(args.obs_shape, args.num_actions, seed=args.seed))
        self.actor_model = (actor_model_fn or (lambda: ActorNet(args.obs_shape, args.num_actions)))()
        self.actor_model.share_memory()                                # :58
        self.ring = self.create_buffers(args.obs_shape, args.num_actions)
        self.buffers = self.ring.buffers
        self.rnn_state_buffers = [tuple() for _ in range(args.num_buffers)]
        args.checkpoint_path = os.path.join(args.output_dir, args.project, args.algo_name)
        os.makedirs(args.checkpoint_path, exist_ok=True)
        self._ctx = mp.get_context('fork')
        self._global_step = self._ctx.Value('q', 0)
        self.learner: Optional[B200ImpalaLearner] = None

    # -------------------------------------------------------------------------------------------------
    @property
    def global_step(self) -> int:
        return int(self._global_step.value)

    def hparams(self) -> ImpalaHParams:
        a = self.args
        return ImpalaHParams(rollout_length=a.rollout_length, batch_size=a.batch_size, num_actions=a.num_actions,
                             discounting=a.discounting, baseline_cost=a.baseline_cost, entropy_cost=a.entropy_cost,
                             reward_clipping=a.reward_clipping, max_grad_norm=a.max_grad_norm, learning_rate=a.learning_rate,
                             alpha=a.alpha, momentum=a.momentum, epsilon=a.epsilon, optimizer=a.optimizer)

    def create_buffers(self, obs_shape, num_actions) -> TrajectoryRing:
        """impala_atari.py:122-151 -- same keys/dtypes/shapes, slot-contiguous shared memory"""
        return TrajectoryRing(self.args.rollout_length, num_actions, self.args.num_buffers, obs_shape)

    # ------------------------------------------------------------------------------------------------- actors
    def get_action(self, actor_index, free_queue, full_queue, actor_model, buffers, rnn_state_buffers) -> None:
        """actor process (impala_atari.py:153-220).  Key collision of the reference (env 'action' vs agent 'action',
        SURVEY.md §0.9) is kept as the reference behaves: the agent's action overwrites the env's."""
        try:
            torch.set_num_threads(1)
            env = self.env_fn()
            env_output = env.reset()
            agent_output = actor_model(env_output)
            while True:
                index = free_queue.get()
                if index is None:
                    break
                for key in env_output:
                    buffers[key][index][0, ...] = env_output[key]
                for key in agent_output:
                    buffers[key][index][0, ...] = agent_output[key]
                for t in range(self.args.rollout_length):
                    with torch.no_grad():
                        agent_output = actor_model(env_output)
                    env_output = env.step(agent_output['action'])
                    for key in env_output:
                        buffers[key][index][t + 1, ...] = env_output[key]
                    for key in agent_output:
                        buffers[key][index][t + 1, ...] = agent_output[key]
                full_queue.put(index)
        except KeyboardInterrupt:
            pass
        except Exception:
            traceback.print_exc()
            raise

    # ------------------------------------------------------------------------------------------------- learner
    def _ensure_learner(self):
        if self.learner is None:
            if not (self.args.use_cuda and torch.cuda.is_available()):
                raise RuntimeError('the B200 ImpalaTrainer needs CUDA (no CPU learner path)')
            self.learner = B200ImpalaLearner(self.hparams(), init_state_dict=self.actor_model.reference_state_dict(), process_group=None)
            self.ring.pin()
            from ...data.feeder import batch_specs, H2D_KEYS
            hp = self.learner.hp
            specs = batch_specs(hp.rollout_length, hp.batch_size, hp.num_actions)
            dev = self.learner.device
            self._copy_stream = torch.cuda.Stream(dev)
            self._dev_batches = [{k: torch.empty(specs[k][0], dtype=specs[k][1], device=dev) for k in H2D_KEYS} for _ in range(2)]
            self._consumed = [None, None]
            self._slot = 0
            # slot-level staging: one H2D copy per trajectory slot (all keys), then one unpack kernel
            self._staging = [torch.empty(hp.batch_size, self.ring.slot_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
            import ctypes
            lay = self.ring.layout
            self._slot_off = (ctypes.c_int64 * 6)(*[lay[k][0] for k in ('obs', 'reward', 'done', 'action', 'policy_logits', 'episode_return')])
            self._publish_host = {n: torch.empty(self.learner.shapes[n], dtype=torch.float32).pin_memory() for n in PARAM_NAMES}

    def get_batch(self, free_queue, full_queue, buffers=None, rnn_state_buffers=None, timings=None, lock=None):
        """impala_atari.py:222-268: dequeue B slot indices, copy them column-wise into a time-major device batch
        (async, pinned, copy stream), release the slots once the copies have finished."""
        self._ensure_learner()
        from ...data.feeder import H2D_KEYS
        buffers = buffers or self.buffers
        if lock is not None:
            with lock:
                indices = [self._dequeue(full_queue) for _ in range(self.args.batch_size)]
        else:
            indices = [self._dequeue(full_queue) for _ in range(self.args.batch_size)]
        s = self._slot
        self._slot ^= 1
        dst = self._dev_batches[s]
        from ... import _lib
        sb = self.ring.slot_bytes
        hp = self.learner.hp
        with torch.cuda.stream(self._copy_stream):
            if self._consumed[s] is not None:
                self._copy_stream.wait_event(self._consumed[s])
            if buffers is self.buffers:       # ring slots: ONE pinned H2D copy per slot, then scatter on the device
                stg = self._staging[s]
                for b, m in enumerate(indices):
                    stg[b].copy_(self.ring.block[m * sb:(m + 1) * sb], non_blocking=True)
                _lib.check(_lib.lib().srl_unpack_slots(
                    stg.data_ptr(), sb, self._slot_off, hp.rollout_length, hp.batch_size, hp.num_actions, dst['obs'].data_ptr(),
                    dst['reward'].data_ptr(), dst['done'].data_ptr(), dst['action'].data_ptr(), dst['policy_logits'].data_ptr(),
                    dst['episode_return'].data_ptr(), self._copy_stream.cuda_stream), 'srl_unpack_slots')
            else:                             # foreign buffer dict (reference-style lists of tensors): per-key column copies
                for b, m in enumerate(indices):
                    for k in H2D_KEYS:
                        dst[k][:, b].copy_(buffers[k][m], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        ev.synchronize()                           # slots are owned until their copy finished
        for m in indices:
            free_queue.put(m)
        torch.cuda.current_stream(self.learner.device).wait_event(ev)
        self._cur_slot = s
        return dst, tuple()

    def _dequeue(self, full_queue):
        """full_queue.get() that notices dead actors: the reference blocks forever when an actor process has died
        (impala_atari.py:238-241); here a crashed actor surfaces as an error in the learner instead of a hang."""
        actors = getattr(self, '_actors', None)
        if not actors:
            return full_queue.get()
        while full_queue.empty():
            dead = [p.name for p in actors if not p.is_alive()]
            if dead:
                raise RuntimeError(f'actor process(es) exited while the learner was waiting for trajectories: {dead}')
            time.sleep(0.0005)
        return full_queue.get()

    def learn(self, actor_model, learner_model, batch, initial_rnn_state=(), lock=None) -> Dict[str, Any]:
        """impala_atari.py:270-349.  ``learner_model`` is ignored (the learner state lives in B200ImpalaLearner)."""
        self._ensure_learner()
        stats = self.learner.learn(batch)
        if not math.isfinite(stats['total_loss']):          # never publish poisoned weights to the actors
            raise FloatingPointError(f'non-finite learner loss: {stats}')
        if os.environ.get('SRL_CHECK_FINITE'):               # debugging aid: name the first poisoned tensors
            bad_p = [n for n, v in self.learner.params.items() if not bool(torch.isfinite(v).all())]
            bad_g = [n for n, v in self.learner.grads.items() if not bool(torch.isfinite(v).all())]
            if bad_p or bad_g:
                raise FloatingPointError(f'step {self.learner.global_step}: non-finite params {bad_p} grads {bad_g} stats {stats}')
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.learner.device))
        if getattr(self, '_cur_slot', None) is not None:
            self._consumed[self._cur_slot] = ev
        self.publish_weights(actor_model)
        return stats

    def publish_weights(self, actor_model) -> None:
        """impala_atari.py:348: actor_model.load_state_dict(learner_model.state_dict()) -- D2H into pinned staging, then
        into the shared-memory actor parameters that the actor processes read lock-free"""
        for n in PARAM_NAMES:
            self._publish_host[n].copy_(self.learner.params[n], non_blocking=True)
        torch.cuda.current_stream(self.learner.device).synchronize()
        actor_model.load_reference_state_dict(self._publish_host)

    def learn_process(self, threading_id, actor_model, learner_model, free_queue, full_queue, buffers, rnn_state_buffers, lock=None):
        """impala_atari.py:351-401"""
        try:
            while self.global_step < self.args.total_steps:
                batch, state = self.get_batch(free_queue, full_queue, buffers, rnn_state_buffers, None, lock)
                stats = self.learn(actor_model, learner_model, batch, state, lock)
                with self._global_step.get_lock():
                    self._global_step.value += self.args.rollout_length * self.args.batch_size   # :391
                self.last_stats = stats
        except KeyboardInterrupt:
            return
        except Exception:
            traceback.print_exc()
            raise

    def train(self, log_every_s: float = 5.0, learner_in_process: bool = True) -> Dict[str, Any]:
        """impala_atari.py:403-494.  Actors are forked BEFORE CUDA is touched; the learner loop then runs in this
        process (learner_in_process) -- one learner per GPU; multi-GPU runs launch one trainer per rank (torchrun)."""
        free_queue, full_queue = self._ctx.SimpleQueue(), self._ctx.SimpleQueue()
        actors = []
        for i in range(self.args.num_actors):
            p = self._ctx.Process(target=self.get_action, name=f'actor-process-{i}',
                                  args=(i, free_queue, full_queue, self.actor_model, self.buffers, self.rnn_state_buffers))
            p.start()
            actors.append(p)
        self._actors = actors
        for m in range(self.args.num_buffers):
            free_queue.put(m)
        timer = timeit.default_timer
        t0, s0 = timer(), self.global_step
        checkpoint_path = os.path.join(self.args.output_dir, self.args.project, 'model.tar')
        try:
            self.learn_process(0, self.actor_model, None, free_queue, full_queue, self.buffers, self.rnn_state_buffers, None)
        finally:
            self._actors = None
            for _ in range(self.args.num_actors):
                free_queue.put(None)
            for p in actors:
                p.join(timeout=2)
                if p.is_alive():
                    p.terminate()
        sps = (self.global_step - s0) / max(timer() - t0, 1e-9)
        self.save_checkpoint(checkpoint_path)
        return dict(steps=self.global_step, sps=sps, **getattr(self, 'last_stats', {}))

    def save_checkpoint(self, checkpoint_path: str) -> None:
        """impala_atari.py:496-515 (same dict keys)"""
        if self.args.disable_checkpoint:
            return
        os.makedirs(os.path.dirname(checkpoint_path), exist_ok=True)
        opt = self.learner.optimizer_state_dict() if self.learner is not None else {}
        torch.save({'model_state_dict': self.actor_model.reference_state_dict(), 'optimizer_state_dict': opt,
                    'hparam': {k: v for k, v in vars(self.args).items()}}, checkpoint_path)
