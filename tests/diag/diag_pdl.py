import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from scalerl_b200 import _lib
L = _lib.lib()
flag = torch.zeros(1, dtype=torch.int32, device='cuda'); out = torch.zeros(512, dtype=torch.int32, device='cuda')
for name, st in (('legacy', None), ('created', torch.cuda.Stream())):
    bad = 0
    for it in range(50):
        out.zero_(); torch.cuda.synchronize()
        _lib.check(L.srl_test_pdl(flag.data_ptr(), out.data_ptr(), 512, 20000, st.cuda_stream if st else None))
        torch.cuda.synchronize()
        bad += int((out != 1).sum())
    print(name, 'blocks that saw flag==0:', bad)
