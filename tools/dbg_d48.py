import sys, os, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import oracle_lib as O
import dvode_b200 as B
from test_gpu_parity import gpu_schedule
nsys=48
par, y0, touts = B.batches.diffusion48_batch(nsys), B.batches.diffusion48_y0(), [1e-4, 1e-3, 1e-2, 1e-1]
rtol, atol, nsplit = B.batches.DIFFUSION48_TOL["rtol"], B.batches.DIFFUSION48_TOL["atol"], 2
for mf,mf2 in ((10,20),(12,-22)):
    ref = O.batch_solve2("diffusion48", par, y0, 0.0, touts, nsplit, rtol, atol, mf, rtol, atol, itol=1, pow_mode=1, mf2=mf2)
    got = gpu_schedule("diffusion48", par, y0, 0.0, touts, rtol, atol, mf, 1, strict=True, per_call=True, rwork_opts=np.zeros(20), iwork_opts=np.zeros(30,dtype=np.int32),
                       phase2=dict(nsplit=nsplit, rtol=rtol, atol=atol, itol=1, itask=1, iopt=0, mf=mf2, rwork_opts=np.zeros(20), iwork_opts=np.zeros(30,dtype=np.int32)))
    eq=[bool(np.array_equal(got["y"][:,k], ref["y"][:,k])) for k in range(4)]
    bad=[i for i in range(nsys) if not np.array_equal(got["y"][i,2], ref["y"][i,2])]
    print(os.environ.get("BVODE_NOSTASH"), mf,mf2,eq, "bad systems", len(bad), bad[:6])
    if bad:
        i=bad[0]
        print(" sys",i,"ref istats@2", ref["istats"][i,2], "got", got["istats"][i,2])
        print(" ref rstats@2", ref["rstats"][i,2], "got", got["rstats"][i,2])
        d=np.abs(got["y"][i,2]-ref["y"][i,2]); print(" maxdiff comp", d.argmax(), d.max(), "ncomp differing", (d>0).sum(), np.nonzero(d>0)[0][:10])
