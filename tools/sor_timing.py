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
# compute warps stamp slots 0 (step start), 2 (operands loaded), 4 (arithmetic done), 5 (stores issued), 6 (barrier passed);
# the producer 0 (start), 1 (bulk copies issued), 2 (stage mbarrier passed), 5 (halo mbarriers passed), 6 (barrier passed)
for wp in range(7):
    slots = [0, 1, 2, 5, 6] if wp == 6 else [0, 2, 4, 5, 6]
    names = ["issue", "stage wait", "halo wait", "barrier"] if wp == 6 else ["loads", "arithmetic", "stores+prefetch", "barrier"]
    d = np.diff(t[wp][:, slots], axis=1).mean(0).round(0)
    role = "producer" if wp == 6 else "k=%d rows %d.." % (wp // 2, 32 * (wp % 2))
    print("warp %d %-14s %s | step %.0f cycles" % (wp, role, dict(zip(names, d)), (t[wp, 1:, 0] - t[wp, :-1, 0]).mean()))
