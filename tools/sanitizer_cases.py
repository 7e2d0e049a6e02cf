"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python tools/sanitizer_cases.py
the smoke configuration (P=8 gray flow) with both exact SOR kernels (sor_lane_kernel: flag-synchronised warps;
sor_wave_kernel: single CTA), a forward-backward case, a P=12 RGB and a P=12 stereo case (window-staged patch
kernel, stereo SOR), a 70-row level as three bands of the lane kernel and forced into a cluster of bands of the
wave kernel with 1 and 2 rows per thread (st.async halo exchange), and the 8-bit frame path (pyramid and
upsampling kernels).  Results are checked against the oracle so that a clean log means a correct run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from of_dis_b200 import api, params, preprocess, synth
from oracle import port_driver

CASES = [
    ("smoke_p8_gray", (128, 256), 1, 2, "3 1 12 12 0.05 0.95 0 8 0.4 0 1 0 1 10 10 5 1 3 1.6 0", {"sor_lane": 1}),
    ("smoke_p8_gray_wave", (128, 256), 1, 2, "3 1 12 12 0.05 0.95 0 8 0.4 0 1 0 1 10 10 5 1 3 1.6 0", {"sor_lane": 0}),
    ("lane_rows70_3bands", (140, 176), 1, 2, "2 1 6 6 0.05 0.95 0 8 0.4 0 1 0 1 10 10 5 1 2 1.6 0", {"sor_lane": 1}),
    ("fbcon_p8_gray", (64, 96), 1, 2, "3 1 8 8 0.05 0.95 0 8 0.4 1 1 0 1 10 10 5 1 3 1.6 0", {}),
    ("p12_rgb_l1", (96, 128), 3, 2, "3 1 8 8 0.05 0.95 0 12 0.75 0 1 1 1 10 10 5 1 3 1.6 0", {}),
    ("p12_stereo", (96, 128), 1, 1, "3 1 8 8 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0", {}),
    ("cluster_rows70_rt1", (140, 176), 1, 2, "2 1 6 6 0.05 0.95 0 8 0.4 0 1 0 1 10 10 5 1 2 1.6 0",
     {"sor_lane": 0, "sor_single_max": 32, "sor_rows_per_thread": 1}),
    ("cluster_rows140_rt2", (140, 96), 1, 1, "1 0 6 6 0.05 0.95 0 8 0.4 0 1 0 1 10 10 5 1 3 1.6 0",
     {"sor_lane": 0, "sor_single_max": 32, "sor_rows_per_thread": 2}),
]
if os.environ.get("SANITIZER_LANE") is not None:  # racecheck: force one SOR kernel everywhere (see profiles/README.md)
    CASES = [(n, sz, ch, nop, num, dict(o, sor_lane=int(os.environ["SANITIZER_LANE"]))) for n, sz, ch, nop, num, o in CASES]
for name, (h, w), ch, nop, numbers, opts in CASES:
    prm = params.from_cli_numbers(numbers.split(), noc=ch, nop=nop)
    i0, i1, _ = synth.synthetic_pair(h, w, ch, seed=5, stereo=(nop == 1), amp=3.0)
    pyr = preprocess.PairPyramids(i0, i1, prm.sc_f, prm.p_samp_s)
    ctx = api.Context(prm, pyr.width, pyr.height, pyr.imgpadding, 2)
    for k, v in opts.items():
        ctx.set_option(k, v)
    frames = np.ascontiguousarray(np.stack([np.stack([i0, i1])] * 2))
    if ch == 1:
        frames = frames[..., None]
    ctx.upload_frames_u8(0, 2, frames, w, h)
    ctx.run(2)
    got = ctx.get_flow(1, prm.sc_l)
    full = np.empty((2, h, w, nop), np.float32)
    ctx.get_flow_fullres(0, 2, full, w, h)
    ctx.sync()
    ctx.close()
    exp = port_driver.port_run(pyr, prm)
    ok = np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    print("%-22s %s" % (name, "bitwise equal to the oracle" if ok else "MISMATCH"), flush=True)
    if not ok:
        sys.exit(1)
print("all cases ok")
