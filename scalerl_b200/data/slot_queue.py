"""SlotQueue -- the free/full index queues of ImpalaTrainer (/root/reference impala_atari.py:416-418: two
``mp.SimpleQueue``s carrying trajectory-slot indices between actor processes and the learner) as a fixed-capacity ring of
int32 in shared memory.

Same interface as ``multiprocessing.SimpleQueue`` for what the trainer uses (``put`` / ``get`` / ``empty``; ``None`` is the
shutdown sentinel of impala_atari.py:181-183), so ``get_action`` / ``get_batch`` work with either.  Why: at B200 learner speed
a step takes ~0.2 ms while one SimpleQueue operation (pickle + pipe write/read + lock) costs ~10-20 us -- 64 of them per
step would make the HOST the bottleneck.  Here an operation is a semaphore and two array accesses, and ``get_many`` takes a
whole batch of indices under one lock acquisition."""
import multiprocessing as mp
from typing import List, Optional

_NONE = -1


class SlotQueue:
    def __init__(self, capacity: int, ctx=None):
        ctx = ctx or mp.get_context('fork')
        self.capacity = int(capacity)
        self._buf = ctx.RawArray('i', self.capacity)         # shared memory, inherited by the forked actors
        self._pos = ctx.RawArray('q', 2)                     # [head (next read), tail (next write)]
        self._lock = ctx.Lock()
        self._items = ctx.Semaphore(0)

    def put(self, index: Optional[int]) -> None:
        v = _NONE if index is None else int(index)
        with self._lock:
            head, tail = self._pos[0], self._pos[1]
            if tail - head >= self.capacity:
                raise RuntimeError('SlotQueue overflow: more indices in flight than slots')
            self._buf[tail % self.capacity] = v
            self._pos[1] = tail + 1
        self._items.release()

    def _pop(self) -> Optional[int]:
        with self._lock:
            head = self._pos[0]
            v = self._buf[head % self.capacity]
            self._pos[0] = head + 1
        return None if v == _NONE else v

    def get(self) -> Optional[int]:
        self._items.acquire()
        return self._pop()

    def get_nowait(self) -> Optional[int]:
        """-> index, or raises IndexError when empty"""
        if not self._items.acquire(block=False):
            raise IndexError('SlotQueue is empty')
        return self._pop()

    def get_many(self, n: int, timeout: Optional[float] = None) -> List[Optional[int]]:
        """up to n indices: blocks (at most ``timeout`` seconds) for the first, takes whatever else is ready"""
        out: List[Optional[int]] = []
        if not self._items.acquire(timeout=timeout):
            return out
        k = 1
        while k < n and self._items.acquire(block=False):
            k += 1
        with self._lock:
            head = self._pos[0]
            for i in range(k):
                v = self._buf[(head + i) % self.capacity]
                out.append(None if v == _NONE else v)
            self._pos[0] = head + k
        return out

    def empty(self) -> bool:
        return self._pos[0] == self._pos[1]

    def qsize(self) -> int:
        return self._pos[1] - self._pos[0]
