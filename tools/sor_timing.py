"""Debug tool: per-phase cycle stamps of the SOR kernel (needs a build with OFDIS_SOR_TIMING=1).
  OFDIS_SOR_TIMING=1 python -m of_dis_b200.build --force && python tools/sor_timing.py"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from of_dis_b200 import api, params, preprocess, synth

prm = params.operating_point(2, 1024)
i0, i1, _ = synth.synthetic_pair(436, 1024, 1, seed=0)
pyr = preprocess.PairPyramids(i0, i1, prm.sc_f, prm.p_samp_s)
ctx = api.Context(prm, pyr.width, pyr.height, pyr.imgpadding, 8)
for f in range(8):
    ctx.upload_pyramids(f, pyr)
ctx.run(8)
ctx.run(8)
ctx.sync()
buf = np.zeros(64 * 8 * 16, np.int64)
assert api.lib().ofdis_debug_sor_times(buf.ctypes.data_as(ctypes.c_void_p)) == 0
t = buf.reshape(64, 8, 16)[:7, :, :7]   # last launch = level 3 (6 compute warps + the producer), steps 40..47
names = ["start", "mbarwait", "lds", "prep", "chain", "stores", "barrier"]
for wp in range(7):
    d = np.diff(t[wp], axis=1)          # per-step phase durations
    role = "producer (issue, -, stage wait, -, -, halo wait, barrier)" if wp == 6 else "k=%d rows %d.." % (wp // 2, 32 * (wp % 2))
    print("warp %d %s: mean cycles per phase %s | step total %.0f" %
          (wp, role, dict(zip(names[1:], d.mean(0).round(0))), (t[wp, 1:, 0] - t[wp, :-1, 0]).mean()))
