import torch, sys
sys.path.insert(0, '.')
from scalerl_b200 import _lib
L = _lib.lib()
g = torch.Generator().manual_seed(0)
res = {}
for mn in (0, 1):
    if mn == 0:
        A = torch.randn(160, 64, generator=g).bfloat16().cuda(); B = torch.randn(64, 64, generator=g).bfloat16().cuda()
    else:
        A = torch.randn(96, 128, generator=g).bfloat16().cuda(); B = torch.randn(96, 64, generator=g).bfloat16().cuda()
    for bo in (0, 1):
        ok = []
        for shift in range(0, 25):
            D = torch.zeros(128, 64, device='cuda')
            _lib.check(L.srl_test_shifted_operand(A.data_ptr(), B.data_ptr(), D.data_ptr(), shift, mn, bo, None))
            torch.cuda.synchronize()
            if mn == 0:
                ref = A[shift:shift + 128].float() @ B.float().t()
            else:
                ref = A[shift:shift + 64].float().t() @ B[shift:shift + 64].float()
            err = (D - ref).abs().max().item() / ref.abs().max().item()
            ok.append(err < 1e-5)
        print('mn_major', mn, 'base_offset_mode', bo, 'ok shifts:', [i for i, o in enumerate(ok) if o], 'bad:', [i for i, o in enumerate(ok) if not o])
