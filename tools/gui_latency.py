import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import modulation_mfcc_b200 as mm
from modulation_mfcc_b200 import _lib
KW = dict(channelN=0, tStep=0.005, winLen=0.025, n_mfcc=13, n_fft=512, minFreq=100, maxFreq=10000, removeFirst=1,
          filtCutoff=12, filtOrd=6, diffMethod="grad", outFilter="iir", outFiltType="low", outFiltCutOff=[12],
          outFiltLen=6, outFiltPolyOrd=3)
y = mm.synth_clip(1, 100_000, 10_000)
for flags in (0, _lib.MMF_FLAG_NO_TC_MEL):
    kw = dict(KW)
    f = lambda: mm.get_MFCCS_change_batch(y[None, :], 10_000, flags=flags, **{k: v for k, v in KW.items() if k != 'channelN'})
    for _ in range(5): f()
    ts = []
    for _ in range(100):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    print('flags', flags, 'median ms', 1e3*np.median(ts), 'min', 1e3*min(ts))
# device-resident single clip, kernel time only
yd = torch.as_tensor(y[None, :]).cuda()
for flags in (0, _lib.MMF_FLAG_NO_TC_MEL):
    f = lambda: mm.get_MFCCS_change_batch(yd, 10_000, flags=flags, **{k: v for k, v in KW.items() if k != 'channelN'})
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): f()
    e1.record(); torch.cuda.synchronize()
    print('device flags', flags, 'ms per call (GPU timeline)', e0.elapsed_time(e1)/50)
