set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests/test_hamming_gpu.py -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import sys, time
sys.path.insert(0,'vi-orb-slam-icra2018_b200')
import torch, orbb200, numpy as np
m = orbb200.Matcher(0)
print('popc peak', m.popc_peak()/1e12, 'Tpopc/s')
P,n=1024,2000
q = torch.randint(0,256,(P,n,32),dtype=torch.uint8,device='cuda'); t = torch.randint(0,256,(P,n,32),dtype=torch.uint8,device='cuda')
qa = torch.rand(P,n,device='cuda')*360; ta = torch.rand(P,n,device='cuda')*360
best=torch.empty(P,n,dtype=torch.int32,device='cuda'); sec=torch.empty_like(best); idx=torch.empty_like(best); m12=torch.empty_like(best); nm=torch.empty(P,dtype=torch.int32,device='cuda')
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for i in range(3): m.bruteforce_device(q,qa,t,ta,0.9,True,best,sec,idx,m12,nm,stream=s.cuda_stream)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for i in range(5): m.bruteforce_device(q,qa,t,ta,0.9,True,best,sec,idx,m12,nm,stream=s.cuda_stream)
    e1.record(s)
s.synchronize()
ms=e0.elapsed_time(e1)/5
print('bruteforce 1024x2000x2000: %.3f ms  %.1f Gcompares/s'%(ms, P*n*n/ms/1e6))
# allpairs
nkf,nd=1024,1000
tab=torch.randint(0,256,(nkf,nd,32),dtype=torch.uint8,device='cuda'); ang=torch.rand(nkf,nd,device='cuda')*360
cnt=torch.empty(256,nkf,dtype=torch.int32,device='cuda')
with torch.cuda.stream(s):
    m.allpairs_device(tab,ang,0,256,0,nkf,0.75,True,cnt,stream=s.cuda_stream)
    e0.record(s)
    m.allpairs_device(tab,ang,0,256,0,nkf,0.75,True,cnt,stream=s.cuda_stream)
    e1.record(s)
s.synchronize()
ms=e0.elapsed_time(e1)
print('allpairs 256x1024 kf x 1000^2: %.3f ms  %.1f Gcompares/s'%(ms, 256*nkf*nd*nd/ms/1e6), cnt.sum().item())
PY
