#!/usr/bin/env python
"""bench.py -- headline benchmark of the trace hot path (BASELINE.json configs[1], "C2"):
AC254-150-AB doublet spot diagram, 2^20 collimated rays (Fibonacci disc, 20 mm), sequential SDF
lens-surface path, Spotdetector at the vendor back focus.  Metric: ray-surface interactions/s.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one full solve_system! of the ray bundle (all waves).  `value` is measured with the rays
resident in HBM (device pointers into the C ABI, spot output left on the device); `e2e` goes through
the same C ABI with pinned HOST buffers, host->device and device->host copies inside the timed
region.  Rank r of an N-GPU run traces its own 2^20-ray shard (weak scaling, no data-path
collective: rays are independent, outputs are disjoint slices).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS = 1 << 20
LAMBDA = 707e-9
WORKLOAD = "C2: AC254-150-AB doublet spot diagram, 2^20 collimated rays, Fibonacci disc d=20mm, Spotdetector at BFL"
CPU_SAMPLE = 1 << 14
# algorithmic FP64 work per unit (SURVEY 8(d) cost table): primitive SDF eval 45, triangle test 40, interaction 60
FLOP_SDF, FLOP_TRI, FLOP_INT = 45.0, 40.0, 60.0
# intersect_wave, 2^20 rays per launch: dram__bytes_read.sum + dram__bytes_write.sum, mean over the 4 waves of one C2 solve
# (ncu --set full, profiles/r01s4_ncu_trace_kernels.txt): 68.1 MB read + 89.0 MB written (the excess over the algorithmic 36 MB of
# writes is register-spill lines evicted from L1)
K1_DRAM_BYTES_PER_LAUNCH = 157.0e6


def rays_for_rank(rank, n=N_RAYS):
    """Fibonacci disc of this rank's shard; ranks get discs rotated by the golden angle so shards differ."""
    from tests import scenes
    pos, d = scenes.fibonacci_disc(n)
    if rank:
        a = rank * 0.1
        c, s = math.cos(a), math.sin(a)
        x, z = pos[:, 0].copy(), pos[:, 2].copy()
        pos[:, 0], pos[:, 2] = c * x - s * z, s * x + c * z
    return np.ascontiguousarray(pos), np.ascontiguousarray(d)


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        mhz = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(mhz)}


def run_reference(args, rank):
    """Reference arm: the reference's CPU algorithm for this path (oracle port: Julia is not installed
    and the reference has no compiled component) on all host cores, bounded sample per step."""
    if rank != 0:
        return
    import __graft_entry__ as ge
    ge.build_oracle()
    from oracle import oracle as orc
    from tests import scenes
    cores = os.cpu_count() or 1
    osc = scenes.doublet_spot_oracle()
    pos, d = rays_for_rank(0, N_RAYS)
    sel = np.linspace(0, N_RAYS - 1, CPU_SAMPLE).astype(np.int64)   # evenly spread over the pupil
    pos, d = np.ascontiguousarray(pos[sel]), np.ascontiguousarray(d[sel])
    times, inter = [], 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = orc.bulk_trace_rays(osc["system"], pos, d, LAMBDA, r_max=100, nthreads=cores, want_segments=False, spot=osc["spot"])
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
            inter = r["interactions"]
    total = sum(times)
    value = inter * len(times) / total
    sample = f"{CPU_SAMPLE} of {N_RAYS} rays per step (evenly spaced over the pupil), threaded driver over rays (BASELINE.md B2/B3)"
    print(json.dumps({
        "impl": "reference", "metric": "ray-surface interactions/s", "value": value, "unit": "interactions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_gpu": N_RAYS, "r_max": 100},
        "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=N_RAYS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-detector", action="store_true", help="skip the secondary Photodetector (C3) measurement")
    ap.add_argument("--c3-side", type=int, default=64, help="beamlets per side of the C3 lattice block (256 = the full 65536-beamlet config)")
    ap.add_argument("--c3-pixels", type=int, default=2048)
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary retrace (a4) and PSFDetector (N2) measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    # stdout carries exactly one JSON line: native libraries that print to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build_libbmo()
    if world > 1:
        dist.barrier()
    m = ge.load_package()
    from bmo_b200 import _lib as L
    from tests import scenes
    dev = local_rank
    n = args.rays
    stream = torch.cuda.current_stream()
    L.set_stream(stream.cuda_stream, dev)

    sc = scenes.doublet_spot(m)
    dsys = m.upload_system(sc["system"], [LAMBDA], device=dev)       # system upload: outside the timed region
    pos_h, dir_h = rays_for_rank(rank, n)
    lam_h = np.zeros(n, dtype=np.int32)
    # pinned host buffers (e2e) and resident device copies (value)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    pos_p, dir_p, lam_p = pin(pos_h), pin(dir_h), pin(lam_h)
    pos_d, dir_d, lam_d = pos_p.cuda(non_blocking=True), dir_p.cuda(non_blocking=True), lam_p.cuda(non_blocking=True)
    spot_obj_p = torch.empty(n, dtype=torch.int32).pin_memory()
    spot_xz_p = torch.empty((n, 2), dtype=torch.float64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    torch.cuda.synchronize()

    def step_device():
        res = m.trace_rays(dsys, (pos_d.data_ptr(), n), dir_d.data_ptr(), lam_d.data_ptr(), None, None, 100,
                           keep_segments=False, device_inputs=True)
        inter = res.interactions
        res.free()
        return inter

    def step_e2e():
        import ctypes as C
        h = C.c_void_p()
        # solve_system! + Spotdetector.data in one C-ABI call: host buffers in, host buffers out
        L.check(L.lib().bmo_trace_rays_spots(dsys.h, n, C.c_void_p(pos_p.data_ptr()), C.c_void_p(dir_p.data_ptr()), C.c_void_p(lam_p.data_ptr()),
                                             None, None, 100, 0, C.c_void_p(spot_obj_p.data_ptr()), C.c_void_p(spot_xz_p.data_ptr()), C.byref(h)))
        info = L.bmo_result_info()
        L.check(L.lib().bmo_result_get_info(h, C.byref(info)))
        L.lib().bmo_result_free(h)
        return info.interactions

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        L.counters_reset(dev)                    # per-kernel counters cover the timed steps only
        times, inter = [], 0
        for _ in range(steps):
            flush.fill_(1)                       # evict inputs from L2 between timed iterations
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            inter = fn()
            e1.record(stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        return times, inter

    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    times, inter = timed(step_device, args.steps, args.warmup)
    c = L.counters(dev)
    n_total_steps = args.steps
    times_e, inter_e = timed(step_e2e, args.steps, args.warmup)
    sampler.stop_flag = True

    def agg(times):
        t = torch.tensor([sum(times)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)    # max over ranks
        return float(t.item())
    tot_ms, tot_ms_e = agg(times), agg(times_e)
    it = torch.tensor([float(inter), float(inter_e)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
    inter_all, inter_all_e = float(it[0].item()), float(it[1].item())

    peak = L.measure_fp64_peak(dev)
    # the retrace leg runs before the detector leg: after the detector's 64 MiB fields have gone through the stream-ordered pool
    # the retrace's wave buffers (4 x 100 MB per call) hit a fragmented pool and its timings scatter between 6 and 30 ms
    extras = None
    if rank == 0 and not args.no_extras:
        extras = {"retrace": bench_retrace(m, L, dev, stream, dsys, sc, pos_d, dir_d, lam_d, n, flush)}
    det = None
    if not args.no_detector:
        det = bench_detector(m, L, dev, stream, args.c3_side, args.c3_pixels, max(2, min(args.steps, 3)), 1, flush, peak, rank, world)
    if extras is not None:
        extras["psf"] = bench_psf(m, L, dev, stream, flush, peak)
    if rank == 0:
        value = inter_all * args.steps / (tot_ms * 1e-3)
        e2e = inter_all_e * args.steps / (tot_ms_e * 1e-3)
        # roofline of the dominant kernel (trace_step, FP64-pipe bound): algorithmic flops per launch / mean launch time
        flops = (FLOP_SDF * c["sdf_evals"] + FLOP_TRI * c["tri_tests"] + FLOP_INT * c["interactions"])
        k1_ms, k1_n = c["trace_step_ms"], max(c["trace_step_launches"], 1)
        achieved = flops / (k1_ms * 1e-3) / 1e12 if k1_ms > 0 else None
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        comp = bench_compaction(m, L, dev, dsys, n, hbm_peak)
        out = {
            "metric": "ray-surface interactions/s", "value": value, "unit": "interactions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_gpu": n, "r_max": 100, "l2": "flushed with a 256 MiB write between timed iterations",
                       "outputs": "spot (x,z) per ray; segment table not kept"},
            "e2e": {"value": e2e, "unit": "interactions/s", "ms_per_step": tot_ms_e / args.steps,
                    "h2d_bytes_per_step": int(pos_h.nbytes + dir_h.nbytes + lam_h.nbytes),
                    "d2h_bytes_per_step": int(spot_obj_p.numel() * 4 + spot_xz_p.numel() * 8)},
            "gpu_launches": int(c["kernel_launches"] * args.steps / n_total_steps),
            "roofline": {"bound": "fp64", "kernel": "intersect_wave", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the 4 waves of one solve, from the
                         # ncu --set full capture summarised in profiles/r01s4_ncu_trace_kernels.txt
                         "traffic": K1_DRAM_BYTES_PER_LAUNCH * n / N_RAYS, "traffic_algorithmic": 100.0 * n,
                         "traffic_source": "profiles/r01s4_ncu_trace_kernels.txt (mean of the 4 waves, 2^20 rays per launch; algorithmic = 64 B ray state read + 36 B hit record written per ray)",
                         "peak_source": "measured here: DFMA probe (bmo_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                         "flops_per_launch": flops / k1_n, "ms_per_launch": k1_ms / k1_n,
                         "share_of_step": k1_ms / n_total_steps / (tot_ms / args.steps) if tot_ms else None,
                         "counted": {"sdf_evals_per_step": c["sdf_evals"] / n_total_steps, "interactions_per_step": c["interactions"] / n_total_steps}},
            "roofline_compaction": comp,
            "clocks": sampler.summary(),
        }
        if det is not None:
            out["detector"] = det
        if extras is not None:
            out.update(extras)
        if not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline()
        print(json.dumps(out), file=json_out)
        json_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


FLOP_PAIR_REF = 760.0   # flop-equivalents per pixel-beamlet pair in the reference's operation sequence
FLOP_PAIR = 226.0       # the same weights (add/mul 1, sqrt/div 8, transcendental 40) over the strength-reduced sequence of
                        # pd_field_fast, counted op by op in DESIGN.md "K4" -- the work the kernel actually has to do


def bench_detector(m, L, dev, stream, k_side, pd_n, steps, warmup, flush, peak_tflops, rank=0, world=1):
    """Metric 2 (Photodetector px-beamlets/s) on the C3 workload: Keplerian expander + Photodetector(40 mm, pd_n),
    k_side^2 beamlets of the C3 lattice (pitch 8 mm / 256, w0 = 1.5 pitch, lambda = 1 um).  The default run uses the central
    64 x 64 block of the 256 x 256 lattice (bounded so that bench.py stays within minutes); the pair rate does not depend on it."""
    import ctypes as C
    import torch
    from tests import scenes2 as s2
    sc = s2.expander(m, pd_n)
    import torch.distributed as dist
    lat = s2.beamlet_lattice(k_side, aperture=8e-3 * k_side / 256)
    pitch = 8e-3 / 256
    lat["pos"][:, 0] += (rank - (world - 1) / 2) * k_side * pitch      # rank r holds the r-th block of the lattice (along x)
    nb = k_side * k_side
    bundle = m.BeamletBundle.from_params(lat["pos"], lat["dir"], lat["lam"], lat["w0"], M2=lat["M2"], P0=1e-3 / 65536, support=lat["support"])
    dsys = m.upload_system(sc["system"], [lat["lam"]], device=dev)
    res = m.trace_beamlets(dsys, bundle.rays, np.zeros(nb, np.int32), bundle.w0, bundle.E0)
    pd_index = dsys.flat.object_index(sc["pd"])
    field_d = torch.zeros(pd_n * pd_n * 2, dtype=torch.float64, device="cuda")
    field_h = torch.zeros(pd_n * pd_n * 2, dtype=torch.float64).pin_memory()

    def run(ptr, flags):
        if flags & L.INPUT_DEVICE:
            field_d.zero_()                      # partial field of this rank
        L.check(L.lib().bmo_pd_accumulate(dsys.h, res.h, pd_index, 0, C.c_void_p(ptr), flags))
        if world > 1 and (flags & L.INPUT_DEVICE):
            dist.all_reduce(field_d, op=dist.ReduceOp.SUM)

    out = {}
    for name, ptr, flags in (("device", field_d.data_ptr(), L.INPUT_DEVICE), ("e2e", field_h.data_ptr(), 0)):
        for _ in range(warmup):
            run(ptr, flags)
        L.counters_reset(dev)
        times = []
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run(ptr, flags)
            e1.record(stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        c = L.counters(dev)
        tot = torch.tensor([sum(times) / steps], dtype=torch.float64, device="cuda")
        pr = torch.tensor([c["px_beamlets"] / steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)       # max over ranks
            dist.all_reduce(pr, op=dist.ReduceOp.SUM)        # whole-job pairs
        out[name] = dict(ms=float(tot.item()), pairs=float(pr.item()), k4_ms=c["pd_field_ms"] / steps, launches=c["kernel_launches"] / steps,
                         pairs_rank=c["px_beamlets"] / steps)
        if name == "device" and world > 1:       # the e2e leg (host field) is a single-GPU notion: rank-local, no collective
            break
    if "e2e" not in out:
        out["e2e"] = dict(out["device"])
    pairs = out["device"]["pairs"]
    ach = FLOP_PAIR * out["device"]["pairs_rank"] / (out["device"]["k4_ms"] * 1e-3) / 1e12
    res.free()
    return {
        "metric": "Photodetector px-beamlets/s", "unit": "px-beamlets/s",
        "value": pairs / (out["device"]["ms"] * 1e-3), "ms_per_step": out["device"]["ms"],
        "e2e": {"value": pairs / (out["e2e"]["ms"] * 1e-3), "ms_per_step": out["e2e"]["ms"],
                "h2d_bytes_per_step": int(field_h.numel() * 8), "d2h_bytes_per_step": int(field_h.numel() * 8)},
        "config": {"workload": f"C3: Keplerian beam expander, {nb} of 65536 GaussianBeamlets per GPU (a {k_side}x{k_side} block of the 256x256 lattice) onto a {pd_n}^2 Photodetector, coherent field sum"
                               + (f", all-reduce of the {pd_n}^2 complex128 field over {world} GPUs inside the timed region" if world > 1 else ""),
                   "beamlets_per_gpu": nb, "pixels": pd_n * pd_n, "pairs_per_step": pairs},
        "n_gpus": world, "scaling": "weak",
        "gpu_launches": out["device"]["launches"],
        "roofline": {"bound": "fp64", "kernel": "pd_field_fast", "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s (FP64 flop-equivalents)",
                     "frac": ach / peak_tflops if peak_tflops else None, "ms_per_launch": out["device"]["k4_ms"],
                     "flop_equiv_per_pair": FLOP_PAIR, "reference_sequence_flop_equiv_per_pair": FLOP_PAIR_REF,
                     "achieved_in_reference_sequence_units": ach * FLOP_PAIR_REF / FLOP_PAIR,
                     "share_of_step": out["device"]["k4_ms"] / out["device"]["ms"]},
    }


def _time_ms(fn, stream, flush, steps=5, warmup=2):
    import torch
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"[bench] {getattr(fn, '__name__', 'step')}: ms per step {['%.3f' % t for t in ts]}", file=sys.stderr)
    return sorted(ts)[len(ts) // 2]      # median of the timed steps


def bench_retrace(m, L, dev, stream, dsys, sc, pos_d, dir_d, lam_d, n, flush):
    """solve_system!(...; retrace=true) of the solved C2 bundle (bmo_retrace: every stored ray
    re-validated against the object it hit before) next to a fresh non-sequential solve that keeps the same outputs
    (segment table kept in both: the retrace needs the stored path)."""
    prev = m.trace_rays(dsys, (pos_d.data_ptr(), n), dir_d.data_ptr(), lam_d.data_ptr(), None, None, 100, keep_segments=True, device_inputs=True)
    prev.keep = True
    inter = {}

    def fresh():
        r = m.trace_rays(dsys, (pos_d.data_ptr(), n), dir_d.data_ptr(), lam_d.data_ptr(), None, None, 100, keep_segments=True, device_inputs=True)
        inter["fresh"] = r.interactions
        r.free()

    def again():
        r = m.retrace(dsys, prev, 100, keep_segments=True)
        inter["retrace"] = r.interactions
        r.free()
    ms_f = _time_ms(fresh, stream, flush)
    L.counters_reset(dev)
    ms_r = _time_ms(again, stream, flush, steps=5, warmup=2)
    c = L.counters(dev)
    prev.free()
    return {"workload": "C2 bundle, 2^20 rays already solved, same poses: bmo_retrace vs a fresh bmo_trace_rays, segment table kept by both, rays resident in HBM",
            "fresh_ms": ms_f, "retrace_ms": ms_r, "interactions": inter.get("retrace"), "interactions_fresh": inter.get("fresh"),
            "retrace_interactions_per_s": inter.get("retrace", 0) / (ms_r * 1e-3),
            "sdf_evals_per_retrace": c["sdf_evals"] / 7.0}


FLOP_PSF_PAIR = 48.0    # per pixel-hit pair: 3 FMA phase, sincos as 40 (the transcendental weight of SURVEY 8(d)), 2 FMA accumulate


def bench_psf(m, L, dev, stream, flush, peak_tflops, n_rays=1 << 16, n_px=512):
    """N2: PSFDetector intensity map (PSFDetector.jl:190-237) of the reference's Airy-disc scene (test/runtests.jl:2765-2802)
    with 2^16 rays on a 512^2 grid: pixel-hit pairs/s of psf_intensity_kernel."""
    import ctypes as C
    import torch
    from tests import scenes
    lens = m.SphericalLens(100e-3, math.inf, 1e-3, 25.4e-3, 1.5)
    psfd = m.PSFDetector(10e-3)
    psfd.translate3d_([0.0, 200e-3 + 0.13e-3, 0.0])
    system = m.System([lens, psfd])
    pos, d = scenes.fibonacci_disc(n_rays, diameter=15e-3, y0=-10e-3)
    res = m.solve_system_(system, m.RayBundle(pos, d, 1e-6))
    hits = len(psfd)
    dsys, oi = psfd._dsys, psfd._index
    lims = np.array(psfd.calc_local_lims(5.0, "bbox"))
    out_d = torch.zeros(n_px * n_px, dtype=torch.float64, device="cuda")
    out_h = torch.zeros(n_px * n_px, dtype=torch.float64).pin_memory()

    def run(ptr, flags):
        L.check(L.lib().bmo_psf_intensity(dsys.h, C.c_void_p(psfd._psf), oi, 0, n_px, L.ptr(lims), 0.0, 0.0, C.c_void_p(ptr), flags))
    ms_d = _time_ms(lambda: run(out_d.data_ptr(), L.INPUT_DEVICE), stream, flush, steps=3, warmup=1)
    k_ms = L.counters(dev)["psf_ms"]
    ms_h = _time_ms(lambda: run(out_h.data_ptr(), 0), stream, flush, steps=3, warmup=1)
    pairs = float(hits) * n_px * n_px
    ach = FLOP_PSF_PAIR * pairs / (k_ms * 1e-3) / 1e12
    psfd.empty_()
    res.free()
    return {"metric": "PSFDetector px-hits/s", "value": pairs / (ms_d * 1e-3), "ms_per_step": ms_d,
            "e2e": {"value": pairs / (ms_h * 1e-3), "ms_per_step": ms_h, "h2d_bytes_per_step": 32, "d2h_bytes_per_step": n_px * n_px * 8},
            "config": {"workload": f"Airy-disc scene of test/runtests.jl:2765-2802, {hits} ray hits, {n_px}^2 pixels, crop_factor 5, :bbox window"},
            "roofline": {"bound": "fp64", "kernel": "psf_intensity_kernel", "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s (FP64 flop-equivalents)",
                         "frac": ach / peak_tflops if peak_tflops else None, "ms_per_launch": k_ms, "flop_equiv_per_pair": FLOP_PSF_PAIR}}


def bench_compaction(m, L, dev, dsys, n, hbm_peak):
    """K3 (queue compaction, HBM-bound) does not run in C2 proper -- every ray lives for all 4 waves.  Overfilling the pupil
    (disc of 40 mm on the 25.4 mm doublet: 60 % of the rays miss everything on the first wave) makes the dead slots the
    majority, which triggers one compaction of the queue; achieved GB/s = algorithmic bytes / CUDA-event time.  Measured on
    the 2^20-slot queue of the headline bundle (75 MB: launch + grid-barrier latency is a third of the time) and on a
    2^23-slot queue (604 MB), where the bytes are what bounds it."""
    import torch
    from tests import scenes

    def one(nq):
        pos, d = scenes.fibonacci_disc(nq, diameter=40e-3)
        pos_d, dir_d = torch.from_numpy(pos).cuda(), torch.from_numpy(d).cuda()
        lam_d = torch.zeros(nq, dtype=torch.int32, device="cuda")
        best = None
        for it in range(4):
            L.counters_reset(dev)
            res = m.trace_rays(dsys, (pos_d.data_ptr(), nq), dir_d.data_ptr(), lam_d.data_ptr(), None, None, 100, keep_segments=False, device_inputs=True)
            res.free()
            c = L.counters(dev)
            if it and c["scatter_ms"] > 0:
                gbs = c["scatter_bytes"] / (c["scatter_ms"] * 1e-3) / 1e9
                best = gbs if best is None else max(best, gbs)
        return best

    best, big = one(n), one(8 * n)
    return {"bound": "hbm", "kernel": "compact_fused (one cooperative launch: count, grid barrier, offsets, scatter)", "achieved": best, "peak": hbm_peak,
            "unit": "GB/s", "frac": (best / hbm_peak) if best else None, "peak_source": "MEASURED_PEAKS.json hbm_gbs",
            "workload": f"C2 doublet with an overfilled pupil (40 mm disc): one compaction of the {n}-slot queue to the ~40 % live rays",
            "large_queue": {"slots": 8 * n, "achieved": big, "frac": (big / hbm_peak) if big else None}}


def cpu_baseline():
    """The oracle port on the host cores over a bounded sample of the same workload (~10-30 s of CPU work)."""
    import __graft_entry__ as ge
    ge.build_oracle()
    from oracle import oracle as orc
    from tests import scenes
    cores = os.cpu_count() or 1
    osc = scenes.doublet_spot_oracle()
    pos, d = rays_for_rank(0, N_RAYS)
    sel = np.linspace(0, N_RAYS - 1, CPU_SAMPLE).astype(np.int64)
    pos, d = np.ascontiguousarray(pos[sel]), np.ascontiguousarray(d[sel])
    t0 = time.perf_counter()
    reps, inter = 0, 0
    while reps < 2 or time.perf_counter() - t0 < 10.0:
        inter += orc.bulk_trace_rays(osc["system"], pos, d, LAMBDA, r_max=100, nthreads=cores, want_segments=False, spot=osc["spot"])["interactions"]
        reps += 1
        if reps >= 8:
            break
    dt = time.perf_counter() - t0
    return {"value": inter / dt, "unit": "interactions/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x {CPU_SAMPLE} of {N_RAYS} rays (evenly spaced over the pupil), OpenMP threaded driver over rays; "
                      "restatement of the Julia reference (Julia is not installed), not the reference itself"}


if __name__ == "__main__":
    main()
