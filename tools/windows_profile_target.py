"""to_windows_triples and to_windows_triples_cbow at 0.5 M walks for ncu (profiles/r02_ncu_extract.txt):

    ncu --set full --clock-control none -k regex:windows_kernel -c 4 -o gpurun_out/r2_prof_win python tools/windows_profile_target.py
"""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_random_walk_b200 import native, rmat
triples = rmat.kg_triples(14541, 237, 310116, device="cuda")
_, ts = rmat.relation_tail_index(triples, 14541)
tw = torch.randint(0, 14541, (500000, 81), dtype=torch.int64, device="cuda")
for _ in range(2):
    out = native.to_windows_triples(tw, 5, 14541, 14778, ts, 1)
    del out
    out = native.to_windows_triples_cbow(tw, 5, 14541, 14778, ts, 1)
    del out
torch.cuda.synchronize()
print("ok")
