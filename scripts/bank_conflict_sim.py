"""Shared-memory bank-conflict model of the lane kernels gather (x_j) / scatter (column-major scratch) accesses,
from the fragment-major records (development aid, CPU only): a 64-bit warp access is served per half warp, one
wavefront per distinct 8-byte word that shares a bank pair.  Baseline order vs a greedy per-lane permutation of the
alignments over the four slots.   python scripts/bank_conflict_sim.py"""
import numpy as np, itertools, sys
sys.path.insert(0,'.')
import layout_fm
from stellarscope_b200 import layout_host, synth
d = synth.generate(n_cells=60, n_alignments=30000, n_loci=28000, seed=1)
L = layout_host.build(d.indptr, d.indices, d.scores, d.row_cell, d.n_cells, d.K)
def waves(words):   # words: list of 8-byte word indices of the active lanes of a half warp
    if not words: return 0
    bins={}
    for w in set(words): bins[w%16]=bins.get(w%16,0)+1
    return max(bins.values())
tot_base=tot_greedy=acc_n=0
for t in range(L.n_tiles):
    s=layout_fm.tile_slices(L,t)
    threads=min(512,max(32,(int(s['hdr'][10])+63)//64*32))
    fm=layout_fm.build_fm(L,t,threads,2,4)
    R=fm.rows; n=(R['meta']&0xff).astype(int); lo=((R['off']&0xffff)>>3).astype(int); hi=(R['off']>>19).astype(int)
    for hw in range(0,len(R),16):
        lanes=[l for l in range(hw,hw+16) if n[l]>0]
        if not lanes: continue
        # baseline
        for p in range(4):
            act=[l for l in lanes if n[l]>p]
            tot_base+=waves([lo[l,p] for l in act])+waves([hi[l,p] for l in act]); 
        acc_n+=sum(n[l] for l in lanes)
        # greedy: choose per lane a permutation of its alignments over the 4 slots (any slot positions)
        gb=[{} for _ in range(4)]; sb=[{} for _ in range(4)]   # per slot: bank -> set(words)
        def cost_add(bins,w):
            b=bins.get(w%16,set()); 
            return len(b|{w})
        for l in lanes:
            k=int(n[l]); best=None
            for slots in itertools.permutations(range(4),k):
                c=0
                for i,p in enumerate(slots):
                    c+=max(0,cost_add(gb[p],lo[l,i])-max([len(v) for v in gb[p].values()],default=0))
                    c+=max(0,cost_add(sb[p],hi[l,i])-max([len(v) for v in sb[p].values()],default=0))
                if best is None or c<best[0]: best=(c,slots)
            for i,p in enumerate(best[1]):
                gb[p].setdefault(lo[l,i]%16,set()).add(lo[l,i]); sb[p].setdefault(hi[l,i]%16,set()).add(hi[l,i])
        for p in range(4):
            tot_greedy+=max([len(v) for v in gb[p].values()],default=0)+max([len(v) for v in sb[p].values()],default=0)
print('alignments',acc_n,'wavefronts per alignment (gather+scatter): baseline %.3f  greedy slot permutation %.3f'%(tot_base/acc_n,tot_greedy/acc_n))
