import sys, time, json; sys.path.insert(0,'.')
import torch
from tum_adlr_deep_reinforcement_learning_b200 import batched as bt
from tum_adlr_deep_reinforcement_learning_b200.config import build_config
res={}
for prec,integ,n in (("f64","rk45",65536),("f64","rk4",65536),("f32","rk4",65536),("f32","rk45",65536),("f64","rk45",4096)):
    cfg=build_config(sim_config_kw={"turbulence":True}, precision=prec, integrator=integ, rk4_substeps=4)
    env=bt.BatchedFixedWing(n,cfg=cfg); env.reset()
    a=torch.rand(n,3,device='cuda')*2-1
    for _ in range(5): env.step(a)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    K=50
    e0.record()
    for _ in range(K): env.step(a)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/K
    nf=env.get_field(bt.FIELD_NFEV).float().mean(0).tolist()
    res["%s_%s_%d"%(prec,integ,n)]=dict(ms=ms, steps_per_s=n/ms*1e3, nfev=nf)
    print(prec,integ,n,"ms/step %.3f  env-steps/s %.3e  mean nfev %s"%(ms,n/ms*1e3,nf), flush=True)
    env.close()
for p in ("f64","f32"):
    res["peak_"+p]=bt.measure_fma_peak(0,p); print("peak",p,res["peak_"+p])
json.dump(res,open('gpurun_out/qb.json','w'),indent=1)
