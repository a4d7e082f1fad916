"""Manual fuzz campaign (not collected by pytest): the C oracle against tests/pyref.py on random snapshots —
whole round, PreFilter, Filter matrix and the pod-at-a-time walk.  `python tests/fuzz_crosscheck.py [n_seeds]`.
Last run: 2500 seeds (4-16 lanes, cases A/B/mixed, 1-70 pods, 1-45 nodes, 1-12 groups, big values every 7th): 0 mismatches."""
import os
import sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import numpy as np, pyref
from randsnap import random_snapshot
from oracle import oracle
bad=0
N_SEEDS = int(sys.argv[1]) if len(sys.argv) > 1 else 400
for seed in range(N_SEEDS):
    L=[4,5,6,9,12,16][seed%6]; case=["mixed","A","B"][seed%3]
    P=int(np.random.default_rng(seed).integers(1,70)); N=int(np.random.default_rng(seed+1).integers(1,45)); G=int(np.random.default_rng(seed+2).integers(1,12))
    vs = "big" if seed%7==0 else "normal"
    snap=random_snapshot(50000+seed,P=P,N=N,G=G,L=L,case=case,value_scale=vs)
    try:
        r=oracle.round(snap,want_bitmap=True,want_score=True,want_filter=True)
        py=pyref.round_outputs(snap)
        bits=np.unpackbits(r.fit_bitmap.view(np.uint8),axis=1,bitorder="little")[:,:N].astype(bool)
        ok=(bits==py["fit"]).all() and (r.score==py["score"]).all()
        for k in ("feasible_count","best_node","best_score","admit","order","rank"): ok = ok and (getattr(r,k)==py[k]).all()
        codes,denied,m=pyref.prefilter_round(snap); ok = ok and (r.prefilter==codes).all() and (r.new_denied==denied).all() and r.max_group==m
        fp,fc=pyref.filter_round(snap); fb=np.unpackbits(r.filter_bitmap.view(np.uint8),axis=1,bitorder="little")[:,:N].astype(bool)
        ok = ok and (fb==fp).all() and (r.filter_code==fc).all()
        q=None if seed%2 else np.random.default_rng(seed).permutation(P)
        pf,node,ready,_=oracle.replay(snap,q); a,b,c=pyref.replay(snap,q)
        ok = ok and (pf==a).all() and (node==b).all() and (ready==c).all()
    except Exception as e:
        ok=False; print("EXC",seed,repr(e)[:200])
    if not ok: bad+=1; print("MISMATCH seed",seed,L,case,P,N,G,vs, flush=True)
print("done bad",bad)
