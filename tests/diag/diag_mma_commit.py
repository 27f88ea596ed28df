"""does a tcgen05.commit between batches of MMAs cost tensor-pipe time?  python tests/diag/diag_mma_commit.py
(srl_test_mma_rate with the commit cadence packed into `shift`: bits 8-15 = commit every n MMAs, bits 16-19 = commits per point)"""
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
    print('mode N every ncommit issuers | issue clk/MMA  total clk/MMA')
    for mode in (0, 1, 2, 3):
        for N in ((64, 128) if mode == 3 else (32, 64, 128)):
            for every, nc in ((0, 0), (16, 1), (16, 2)):
                for issuers in (1, 2):
                    out.zero_()
                    for _ in range(2):
                        _lib.check_hook(H.srl_test_mma_rate(N, 21 | (every << 8) | (nc << 16) | (mode << 20), reps, issuers, out.data_ptr(), None))
                    torch.cuda.synchronize()
                    o = out.cpu().tolist()
                    s = '  '.join(f'w{w}: {o[2 * w] / reps:6.1f} {o[2 * w + 1] / reps:6.1f}' for w in range(issuers))
                    print(f'{mode} {N:3d} {every:3d} {nc} {issuers} | {s}')


if __name__ == '__main__':
    main()
