"""tcgen05.mma issue / completion cost from one thread (and two) -- srl_test_mma_rate.  python tests/diag/diag_mma_rate.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch                                    # noqa: E402
from scalerl_b200 import _lib                   # noqa: E402


def main():
    H = _lib.hooks()
    out = torch.zeros(4, dtype=torch.int64, device='cuda')
    reps = 256
    print('N shift issuers | issue clk/MMA  total clk/MMA  (floor 128*N/256)')
    for N in (32, 64, 128, 256):
        for shift in (0, 1, 8, 21):
            for issuers in (1, 2):
                out.zero_()
                for _ in range(2):
                    _lib.check_hook(H.srl_test_mma_rate(N, shift, reps, issuers, out.data_ptr(), None))
                torch.cuda.synchronize()
                o = out.cpu().tolist()
                s = '  '.join(f'w{w}: {o[2 * w] / reps:6.1f} {o[2 * w + 1] / reps:6.1f}' for w in range(issuers))
                print(f'{N:3d} {shift:3d} {issuers} | {s}   ({128 * N // 256})')


if __name__ == '__main__':
    main()
