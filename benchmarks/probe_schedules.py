import sys, os, json, numpy as np, torch
sys.path.insert(0, '/root/repo')
from deepfmkit_b200 import _lib
from deepfmkit_b200 import fit as tun
ctx=_lib.Context(0); ctx.use_torch_stream()
C,T,R,N=256,int(100*200e3),4000,10
x=torch.empty((C,T),dtype=torch.float64,device='cuda')
ctx.synth_snr_dev(x.data_ptr(),T,C,200e3,1000.0,6.0,dphi=2*np.pi/C,seed=5)
rows=torch.empty((C,T//R,8),dtype=torch.float64,device='cuda')
opts=tun.current_lm_opts(); w0=2*np.pi*1000/200e3
for sched in (True, 16):
    for rep in range(3):
        ctx.profile_enable(True); ctx.profile_read(reset=True)
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); ctx.nls_fit_batch_dev(x.data_ptr(),C,T//R,T,R,N,w0,[1.6,6.0,0,0],None,0,sched,opts,rows.data_ptr()); b.record(); b.synchronize()
        p=ctx.profile_read(reset=True)
    a2,b2=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a2.record()
    for rep in range(10):
        ctx.nls_fit_batch_dev(x.data_ptr(),C,T//R,T,R,N,w0,[1.6,6.0,0,0],None,0,sched,opts,rows.data_ptr())
    b2.record(); b2.synchronize()
    back_to_back=a2.elapsed_time(b2)/10
    fl=rows[:,:,6].flatten()
    print(json.dumps({"sched":str(sched),"ms":a.elapsed_time(b),"back_to_back_ms":back_to_back,"prof":{k:round(v,3) for k,v in p.items() if k.endswith('_ms')},"flags":{int(k):int((fl==k).sum()) for k in (0,1,2)}}),flush=True)
