"""Phase timeline of CTA 0 of enc_fused_fwd_kernel (SRL_FUSED_DEBUG=1): where do the ~10 us per frame go?
    SRL_FUSED_DEBUG=1 python tests/diag/diag_fused.py [T B]"""
import os
import sys

os.environ['SRL_FUSED_DEBUG'] = '1'
os.environ['SRL_FUSED_FWD'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch                                    # noqa: E402
from oracle import impala_oracle as O           # noqa: E402
from scalerl_b200.learner import B200ImpalaLearner, ImpalaHParams   # noqa: E402


def main():
    T, B = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (20, 32)
    A = 6
    L = B200ImpalaLearner(ImpalaHParams(rollout_length=T, batch_size=B, num_actions=A), init_state_dict=O.init_params(A, seed=0), process_group=False,
                          use_graph=False)
    batch = {k: v.cuda() for k, v in O.synthetic_batch(T, B, A, seed=0).items()}
    for _ in range(3):
        L.forward(batch)
    torch.cuda.synchronize()
    raw = L.debug_buffer('fused_dbg').view(torch.int64).view(5, 8, 8).cpu()
    t0 = int(raw[0, 0, 0])
    rel = lambda v: (int(v) - t0) / 1000.0 if int(v) else float('nan')
    print(f'kernel: start 0.0  after pdl_wait {rel(raw[0,0,1]):.2f}  weights+sync {rel(raw[0,0,2]):.2f}  end {rel(raw[0,0,3]):.2f}  (us)')
    names = {1: ('producer', ['u8 issued']), 2: ('mma', ['c1 t0', 'c1 t1', 'c1 t2', 'c1 t3', 'a1_full ok', 'conv2 issued']),
             3: ('converter', ['u8 ready', 'tile0', 'tile1', 'tile2', 'tile3']),
             4: ('epilogue', ['acc1[0]', 'acc1[1]', 'acc1[2]', 'acc1[3]', 'a1_full sent', 'acc2 ready', 'frame done'])}
    for it in range(6):
        print(f'--- frame {it}')
        for role, (nm, evs) in names.items():
            print(f'  {nm:10s} ' + '  '.join(f'{e} {rel(raw[role, it, i]):.2f}' for i, e in enumerate(evs)))


if __name__ == '__main__':
    main()
