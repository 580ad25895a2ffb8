import time, numpy as np, sys
sys.path.insert(0,'/root/repo')
from aircraftoptimalcontrol_b200.lqr_tracking import lqr_tracking_batch
from tests.util import golden
d=golden("lqr_tracking.npz")
rng=np.random.default_rng(1234)
delta=rng.uniform(-0.1,0.1,(4096,6)); delta[0]=0.1
for rep in range(3):
    t=time.perf_counter()
    xr,ur=lqr_tracking_batch(d["xx_opt"],d["uu_opt"],delta)[:2]
    print("lqr_tracking_batch 4096:", round((time.perf_counter()-t)*1e3,1),"ms")
