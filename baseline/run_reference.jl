# baseline/run_reference.jl -- the UNMODIFIED BeamletOptics.jl reference timed on the host CPUs (BASELINE.md section 4, B1 / B2).
#
#     julia -t $(nproc) --project=/root/reference baseline/run_reference.jl c2 [n_rays]
#     julia -t $(nproc) --project=/root/reference baseline/run_reference.jl c3 [n_beamlets_per_side] [pixels]
#
# B1 "stock": solve_system!(system, source) as the package ships it (the loop over beams(bg) is serial, System.jl:463-468;
#     the Photodetector accumulation is threaded over pixel rows, Photodetector.jl:86-105).
# B2 "threaded driver": Threads.@threads over chunks of beams(bg), one deep-copied System per task (Spotdetector.data push!
#     and Photodetector.field += are not thread-safe), results merged afterwards -- the fair multi-threaded comparison.
# Every config runs once untimed (JIT), then is timed on fresh beams.  One JSON line per arm on stdout.
#
# NOT EXECUTED IN THE BUILD IMAGE (no Julia there); bench.py --impl reference falls back to the C++ restatement (B3).
using BeamletOptics, LinearAlgebra, Printf
const BMO = BeamletOptics

function fibonacci_disc(n, diameter, y0)          # UniformDiscSource formula (BeamGroups.jl:232-243), basis e1 = x, e2 = z
    golden = π * (3 - sqrt(5))
    [(sqrt((i - 0.5) / n) * diameter / 2 * cos(i * golden), y0, sqrt((i - 0.5) / n) * diameter / 2 * sin(i * golden)) for i in 1:n]
end

function c2_system()
    dl = SphericalDoubletLens(87.9e-3, -105.6e-3, Inf, 6e-3, 3e-3, 25.4e-3, 1.6456, 1.7168)     # AC254-150-AB at 707 nm (runtests.jl:1275-1282)
    sd = Spotdetector(5e-3)
    translate3d!(sd, [0, BMO.thickness(dl) + 0.14368, 0])
    return System([dl, sd]), sd
end
c2_beams(n) = [Beam(Ray([p...], [0.0, 1.0, 0.0], 707e-9)) for p in fibonacci_disc(n, 20e-3, -0.05)]

count_interactions(beams) = sum(b -> count(r -> !isnothing(BMO.intersection(r)), BMO.rays(b)), beams)

function run_c2(n)
    sys, sd = c2_system()
    solve_system!(sys, c2_beams(min(n, 256)))                       # JIT
    empty!(sd)
    beams = c2_beams(n)
    t1 = @elapsed for b in beams; solve_system!(sys, b); end         # B1: what solve_system!(system, ::AbstractBeamGroup) does
    I = count_interactions(beams)
    emit("C2", "stock", n, I, t1)
    beams = c2_beams(n)
    nt = Threads.nthreads()
    chunks = collect(Iterators.partition(1:n, cld(n, nt)))
    systems = [c2_system() for _ in chunks]                         # one System + Spotdetector per task
    t2 = @elapsed Threads.@threads for k in eachindex(chunks)
        for i in chunks[k]; solve_system!(systems[k][1], beams[i]); end
    end
    emit("C2", "threaded driver", n, count_interactions(beams), t2)
end

function emit(cfg, arm, n, units, secs; unit = "interactions/s")
    @printf("{\"impl\": \"reference\", \"kind\": \"julia\", \"config\": \"%s\", \"arm\": \"%s\", \"n\": %d, \"value\": %.6e, \"unit\": \"%s\", \"seconds\": %.4f, \"nthreads\": %d, \"cpu_threads\": %d, \"cpu\": \"%s\", \"julia\": \"%s\"}\n",
        cfg, arm, n, units / secs, unit, secs, Threads.nthreads(), Sys.CPU_THREADS, Sys.cpu_info()[1].model, string(VERSION))
end

function c3_system(pixels)
    # docs/src/tutorials/expander.md:25-66: two thin lenses (n = 1.5), second one rotated by 180 deg, spacing f1 + f2; Photodetector(40 mm, pixels) at y = 0.25
    l1 = ThinLens(50e-3, 50e-3, 25.4e-3, 1.5)
    l2 = ThinLens(100e-3, 100e-3, 50.8e-3, 1.5)
    f1 = BMO.lensmakers_eq(50e-3, -50e-3, 1.5); f2 = BMO.lensmakers_eq(100e-3, -100e-3, 1.5)
    translate3d!(l2, [0, f1 + f2, 0])
    pd = Photodetector(40e-3, pixels)
    translate3d!(pd, [0, 0.25, 0])
    return System([l1, l2, pd]), pd
end
function c3_beamlets(k)
    pitch = 8e-3 * k / 256 / k
    c = ((0:k-1) .- (k - 1) / 2) .* pitch
    [GaussianBeamlet([x, -0.05, z], [0.0, 1.0, 0.0], 1e-6, 1.5 * pitch; M2 = 1, P0 = 1e-3 / 65536, support = [1.0, 0.0, 0.0]) for x in c for z in c]
end
function run_c3(k, pixels)
    sys, pd = c3_system(pixels)
    solve_system!(sys, c3_beamlets(1)[1]); empty!(pd)               # JIT
    gs = c3_beamlets(k)
    t = @elapsed for g in gs; solve_system!(sys, g); end             # B1: the field accumulation is already threaded over pixel rows
    emit("C3", "stock", length(gs), length(gs) * pixels^2, t; unit = "px-beamlets/s")
end

cfg = length(ARGS) >= 1 ? lowercase(ARGS[1]) : "c2"
if cfg == "c2"
    run_c2(length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 2^14)
elseif cfg == "c3"
    run_c3(length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 8, length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 2048)
else
    error("unknown config $cfg (c2 | c3)")
end
