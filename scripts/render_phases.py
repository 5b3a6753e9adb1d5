"""Where a render CTA's time goes: per-phase SM clock cycles of thread 0, averaged over CTAs (instrumented build).
    python -m multi_car_racing_b200.build --phase-clocks && MCR_LIB_PATH=multi_car_racing_b200/libmcr_clk.so python scripts/render_phases.py [B] [A]
Thread 0 sits in warp 0, so a phase's figure is warp 0's work plus its wait at the barrier that ends the phase."""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import multi_car_racing_b200 as mcr
from multi_car_racing_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
A = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L = _lib.load()
np.random.seed(1234)
venv = mcr.BatchedMultiCarRacing(B, num_agents=A, auto_reset=False, max_episode_steps=0, seed=1234)
venv.reset(device_tracks=True)
g = torch.Generator(device=venv.device); g.manual_seed(1234)
tape = torch.rand((128, B, A, 3), device=venv.device, generator=g); tape[..., 0] = tape[..., 0] * 2 - 1
names = ["0 tables+pdl wait+hdr", "1 -", "2 -", "3 list -> smem", "4 -", "5 zero masks+barrier", "6 spans", "7 fill", "8 -"]   # fill_kernel (MCR_RENDER_FUSED=1: the fused kernel's phases 0-8)
done = 0
buf = (ctypes.c_ulonglong * 16)()
for name, upto in (("t=0.02", 0), ("t=0.5", 24), ("mid(170)", 170)):
    while done < upto:
        venv.step(tape[done % 128]); done += 1
    torch.cuda.synchronize()
    L.mcr_debug_phase_clocks(None, 1)
    for _ in range(5): venv.render_only()
    torch.cuda.synchronize()
    L.mcr_debug_phase_clocks(buf, 1)
    v = np.array(list(buf), np.float64); n = max(v[15], 1)
    print("%s: %d CTAs, %.0f cycles per CTA" % (name, n, v[:9].sum() / n))
    for k, nm in enumerate(names):
        print("   %-20s %8.0f cycles  %5.1f %%" % (nm, v[k] / n, 100 * v[k] / max(v[:9].sum(), 1)))
    npj = max(v[14], 1)
    print("   project_kernel, warp 0 of %d CTAs: pdl wait+ids %.0f | loads+cull+barrier %.0f | passes: candidates+scan %.0f, prefix hand-off %.0f, edges+stores %.0f  (cycles per CTA)"
          % (npj, v[9] / npj, v[10] / npj, v[11] / npj, v[12] / npj, v[13] / npj))
