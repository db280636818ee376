import sys, numpy as np
sys.path.insert(0,'.')
import spectrogram_rs_b200 as sbx
from tests import signals
for (n,hop) in ((2048,2048),(2400,58),(4096,512)):
    plan = sbx.Plan(sample_rate=48000.0, window_n=n, pad_factor=2, hop=hop, rows_h=1024, channels=2, lut="magma")
    live = plan.live(8192)
    lr = signals.stereo_loopback(n*4)
    live.push(lr[:n])
    live.measure(lr[:hop], 1000)
    lat = live.measure(lr[hop:2*hop], 5000)
    col = live.pop_column() if live.push(lr[:n]) else None
    print(n, hop, 'p50 %.1f p99 %.1f us' % (np.percentile(lat,50), np.percentile(lat,99)))
    live.close(); plan.close()
