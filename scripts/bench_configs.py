#!/usr/bin/env python
"""Timing of BASELINE.json's configs 3-5 at their full sizes on one B200 (bench.py carries config 2 and a
bounded block of config 3).  One JSON line per config; device time = bmo_counters' CUDA-event times of the
C-ABI calls, wall time = the whole call with host buffers (copies included).

    python scripts/bench_configs.py [--c3-side 256] [--c4-rays 10000000] [--c5-poses 4096]
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-side", type=int, default=256)
    ap.add_argument("--c3-pixels", type=int, default=2048)
    ap.add_argument("--c4-rays", type=int, default=10_000_000)
    ap.add_argument("--c5-poses", type=int, default=4096)
    ap.add_argument("--c5-pixels", type=int, default=256)
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build_libbmo()
    m = ge.load_package()
    from bmo_b200 import _lib as L
    from tests import scenes2 as s2

    # ---- C3: Keplerian expander, k^2 GaussianBeamlets onto a 2048^2 Photodetector --------------------
    k, n = args.c3_side, args.c3_pixels
    sc = s2.expander(m, n)
    lat = s2.beamlet_lattice(k, aperture=8e-3 * k / 256)
    bundle = m.BeamletBundle.from_params(lat["pos"], lat["dir"], lat["lam"], lat["w0"], M2=lat["M2"], P0=1e-3 / 65536, support=lat["support"])
    L.counters_reset()
    t0 = time.perf_counter()
    res = m.solve_system_(sc["system"], bundle)
    wall = time.perf_counter() - t0
    c = L.counters()
    pairs = c["px_beamlets"]
    print(json.dumps({"config": f"C3: beam expander, {k * k} GaussianBeamlets onto a {n}^2 Photodetector (coherent field sum)",
                      "beamlets": k * k, "pixels": n * n, "px_beamlets": pairs, "trace_interactions": res.interactions,
                      "trace_ms": c["trace_ms"], "pd_ms": c["pd_ms"], "pd_field_kernel_ms": c["pd_field_ms"],
                      "px_beamlets_per_s_kernel": pairs / (c["pd_field_ms"] * 1e-3), "px_beamlets_per_s_call": pairs / (c["pd_ms"] * 1e-3),
                      "interactions_per_s": res.interactions / (c["trace_ms"] * 1e-3), "wall_s_solve_system": wall,
                      "optical_power_W": sc["pd"].optical_power()}), flush=True)
    res.free()
    del sc, bundle

    # ---- C4: non-sequential mesh scene, PolarizedRays with beamsplitter branching ---------------------
    import torch
    nr = args.c4_rays
    sc = s2.mesh_scene(m)
    pos, d, E0 = s2.jittered_lattice(nr)
    dsys = m.upload_system(sc["system"], [1e-6])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    Ere = np.stack([np.broadcast_to(E0, (nr, 3)).real, np.broadcast_to(E0, (nr, 3)).imag], axis=-1).reshape(nr, 6)
    pos_p, dir_p, e_p, lam_p = pin(pos), pin(np.broadcast_to(d, (nr, 3))), pin(Ere), pin(np.zeros(nr, np.int32))
    pos_d, dir_d, e_d, lam_d = (x.cuda() for x in (pos_p, dir_p, e_p, lam_p))
    torch.cuda.synchronize()
    import ctypes as C

    def call(ptrs, flags):
        h = C.c_void_p()
        L.check(L.lib().bmo_trace_rays(dsys.h, nr, C.c_void_p(ptrs[0].data_ptr()), C.c_void_p(ptrs[1].data_ptr()), C.c_void_p(ptrs[3].data_ptr()),
                                       C.c_void_p(ptrs[2].data_ptr()), None, 100, flags, C.byref(h)))
        info = L.bmo_result_info()
        L.check(L.lib().bmo_result_get_info(h, C.byref(info)))
        L.lib().bmo_result_free(h)
        return info
    for name, ptrs, flags in (("rays resident in HBM, beam table only", (pos_d, dir_d, e_d, lam_d), L.INPUT_DEVICE),
                              ("rays resident in HBM, segment table kept", (pos_d, dir_d, e_d, lam_d), L.INPUT_DEVICE | L.KEEP_SEGMENTS),
                              ("pinned host inputs (100 B/ray copied in), beam table only", (pos_p, dir_p, e_p, lam_p), 0)):
        call(ptrs, flags); call(ptrs, flags)    # warm-up: the stream-ordered pool grows to this call's footprint, the system learns its high-water marks
        L.counters_reset()
        t0 = time.perf_counter()
        info = call(ptrs, flags)
        wall = time.perf_counter() - t0
        c = L.counters()
        print(json.dumps({"config": f"C4: rhomb prism + thin beamsplitter + retroreflector + 17k-triangle mesh mirror (BVH), {nr} PolarizedRays, r_max=100; " + name,
                          "rays": nr, "beams": info.n_beams, "interactions": info.interactions, "waves": info.waves, "tri_tests": c["tri_tests"],
                          "trace_ms": c["trace_ms"], "k1_ms": c["trace_step_ms"], "interactions_per_s": info.interactions / (c["trace_ms"] * 1e-3),
                          "wall_s_call": wall, "kernel_launches": c["kernel_launches"]}), flush=True)
    del dsys, sc, pos_d, dir_d, e_d, lam_d
    torch.cuda.empty_cache()

    # ---- C5: Mach-Zehnder mirror-displacement sweep, all poses in one batch ----------------------------
    P, n = args.c5_poses, args.c5_pixels
    sc = s2.mzi(m, pd_n=n)
    B = s2.MZI_BEAM
    g = m.GaussianBeamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B["M2"], P0=B["P0"], support=B["support"])
    base = sc["m1"].position()

    def apply_pose(p):
        sc["m1"].translate_to3d_(base)
        sc["m1"].translate3d_(s2.mzi_shift(p, P))
    shifts = np.array([s2.mzi_shift(p, P) for p in range(P)])
    for rep in range(2):        # the second pass is the steady state (pools grown, library paged in)
        L.counters_reset()
        t0 = time.perf_counter()
        dev = m.solve_pose_sweep_device(sc["system"], g, P, lambda prog: prog.translate3d_(sc["m1"], shifts), sc["pd"], want_fields=False)
        wall_dev = time.perf_counter() - t0
        c_dev = L.counters()
    L.counters_reset()
    t0 = time.perf_counter()
    out = m.solve_pose_sweep(sc["system"], g, P, apply_pose, sc["pd"], want_fields=False)
    wall = time.perf_counter() - t0
    sc["m1"].translate_to3d_(base)
    c = L.counters()
    pw = dev["power"]
    print(json.dumps({"config": f"C5: Mach-Zehnder, {P} kinematic poses of one mirror batched, one {n}^2 interferogram + optical_power per pose",
                      "poses": P, "px_beamlets": c_dev["px_beamlets"], "interactions": dev["result"].interactions, "trace_ms": c_dev["trace_ms"],
                      "pd_ms": c_dev["pd_ms"], "pd_field_kernel_ms": c_dev["pd_field_ms"], "px_beamlets_per_s_kernel": c_dev["px_beamlets"] / (c_dev["pd_field_ms"] * 1e-3),
                      "wall_s_sweep_poses_on_device": wall_dev, "wall_s_sweep_poses_flattened_on_host": wall,
                      "power_identical_to_host_flattened_sweep": bool(np.array_equal(pw, out["power"])),
                      "power_min_W": float(pw.min()), "power_max_W": float(pw.max())}), flush=True)


if __name__ == "__main__":
    main()
