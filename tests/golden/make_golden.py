#!/usr/bin/env python
"""Writes the golden fixtures of tests/golden/ with the CPU oracle (oracle/, the restatement of the reference pinned by
the reference's own known-answer tests -- Julia is not installed here, so these are oracle outputs, not outputs of the
reference itself; they freeze the oracle so that a later change of it, or of the CUDA path, shows up as a diff).

    python tests/golden/make_golden.py        # from the repository root

c2_doublet_64.npz     AC254-150-AB doublet spot diagram (BASELINE config 2), 64 rays of the Fibonacci disc:
                      per ray up to 8 segment rows [pos(3) dir(3) n t nrm(3) obj] + Spotdetector (x, z)
c1_michelson_32.npz   compact Michelson (BASELINE config 1), one GaussianBeamlet, Photodetector 32 x 32: complex field + power
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as orc   # noqa: E402
from tests import scenes           # noqa: E402


def main():
    osc = scenes.doublet_spot_oracle()
    pos, d = scenes.fibonacci_disc(64)
    ref = orc.bulk_trace_rays(osc["system"], pos, d, 707e-9, max_seg=8, spot=osc["spot"])
    np.savez(os.path.join(HERE, "c2_doublet_64.npz"), pos=pos, dir=d, nseg=ref["nseg"], seg=ref["seg"], spot=ref["spot"],
             interactions=np.int64(ref["interactions"]))
    om = scenes.michelson_oracle(pd_n=32)
    B = scenes.MICHELSON_BEAM
    og = orc.gaussian_beamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B["M2"], support=B["support"])
    orc.solve_system_(om["system"], og)
    np.savez(os.path.join(HERE, "c1_michelson_32.npz"), field=om["pd"].pd_field(32), power=np.float64(om["pd"].pd_power()))
    print("written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
