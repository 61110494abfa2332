import numpy as np, sys
sys.path.insert(0,'/root/repo')
from minicv_b200 import estimation as est, synth
from oracle import ransac
a,b,gt,H=synth.homography_pair(7,600)
tr=ransac.RansacTrace()
Ho,mo=ransac.find_homography(a,b,3.0,200,0.995,fixed_iters=True,trace=tr)
for refine in (False,True):
    r=est.ransac_batch(est.MODEL_H,a,b,[0,600],3.0,200,0.995,stop_mode=est.STOP_FIXED,debug=True,refine=refine)
    print("refine",refine,"ok",r.ok,"best",r.best,"inl",r.inliers,"iters",r.iters_run)
    print("samples eq",(r.dbg_samples[0]==np.array(tr.samples)).all())
    oc=np.array([c[0] if c else -1 for c in tr.counts]); gc=r.dbg_counts[0,:,0]
    print("counts eq",(oc==gc).mean(), oc[:10], gc[:10], "oracle best",tr.best_iter, oc.max(), gc.max())
    print("model0 rel", np.abs(r.dbg_models[0,0,0]-tr.models[0][0]).max())
    print(r.models[0]); print(Ho)
