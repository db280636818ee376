import sys, numpy as np
sys.path.insert(0,'.')
import torch
import spectrogram_rs_b200 as sbx
from tests import signals
n,hop=2400,58
plan = sbx.Plan(sample_rate=48000.0, window_n=n, pad_factor=2, hop=hop, rows_h=1024, channels=2, lut="magma")
live = plan.live(8192)
lr = signals.stereo_loopback(n*4)
live.push(lr[:n]); live.measure(lr[:hop], 1000)
lat = live.measure(lr[hop:2*hop], 5000)
print('live N=2400 p50 %.1f p99 %.1f us' % (np.percentile(lat,50), np.percentile(lat,99)))
live.close()
F=65536; ns=(F-1)*hop+n
x=torch.from_numpy(signals.stereo_loopback(ns)).cuda(); out=torch.empty((F,1024,4),dtype=torch.uint8,device='cuda')
s=torch.cuda.Stream(); torch.cuda.set_stream(s); plan.set_stream(s.cuda_stream)
for _ in range(2): plan.process_batch_device(x,ns,out,F)
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record(s)
for _ in range(3): plan.process_batch_device(x,ns,out,F)
b.record(s); torch.cuda.synchronize()
print('batch N=2400 stereo: %.2f Mframes/s' % (3*F/(a.elapsed_time(b)*1e-3)/1e6))
