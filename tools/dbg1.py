import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import oracle_lib as O, dvode_b200 as B
from test_gpu_parity import gpu_schedule, weighted_err, ROB, rob_touts
nsys=4096
par=B.batches.robertson_batch(nsys); tol=B.batches.ROBERTSON_TOL_REF
ref=O.batch_solve("robertson",par,ROB["y0"],0.0,rob_touts(),tol["rtol"],tol["atol"],21,itol=2,pow_mode=0)
for strict in (True, False):
    got=gpu_schedule("robertson",par,ROB["y0"],0.0,rob_touts(),tol["rtol"],tol["atol"],21,2,strict=strict)
    w=weighted_err(got["y"],ref["y"],tol["rtol"],tol["atol"])
    idx=np.unravel_index(np.argmax(w),w.shape); print(strict, w.max(), idx, got["y"][idx[0]], ref["y"][idx[0]], got["istats"][idx[0],-1], ref["istats"][idx[0],-1])
    print("n>10:", (w.max(axis=(1,2))>10).sum(), "n>1:", (w.max(axis=(1,2))>1).sum(), "median max", np.median(w.max(axis=(1,2))))
    for name in ("nst","nfe","nje","nlu","nni","netf"):
        a=got["istats"][:,-1,O.ISTAT[name]].astype(np.int64).sum(); b=ref["istats"][:,-1,O.ISTAT[name]].astype(np.int64).sum()
        print(name,a,b,(a-b)/b)
