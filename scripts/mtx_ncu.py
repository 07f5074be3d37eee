"""Two ss_mtx_measure / ss_mtx_write passes over 20 M random count triples (for ncu; development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stellarscope_b200.engine import Engine
n = 20_000_000
g = torch.Generator(device='cuda').manual_seed(0)
cell = torch.sort(torch.randint(0, 100000, (n,), device='cuda', generator=g, dtype=torch.int32)).values
locus = torch.randint(0, 28001, (n,), device='cuda', generator=g, dtype=torch.int32)
val = torch.randint(1, 60, (n,), device='cuda', generator=g).double()
eng = Engine(0)
for rep in range(2):
    t = eng.mtx_text(locus, cell, val, True)
torch.cuda.synchronize()
print('bytes', t.numel())
