# baseline/dump_fixtures.jl -- reference outputs of C1 (Michelson) and C2 (doublet) as flat binary files that
# tests/golden/load_reference_fixture.py reads, so that fixtures made by the REAL reference can replace the oracle-generated
# goldens the first time a Julia is available:
#
#     julia --project=/root/reference baseline/dump_fixtures.jl tests/golden/ref
#
# Format of every file: Int64 rank, Int64 dims..., then Float64 data in column-major order (complex: re, im interleaved).
# NOT EXECUTED IN THE BUILD IMAGE (no Julia there).
using BeamletOptics, LinearAlgebra
const BMO = BeamletOptics
outdir = length(ARGS) >= 1 ? ARGS[1] : "tests/golden/ref"
mkpath(outdir)
function dump(name, A::AbstractArray)
    open(joinpath(outdir, name * ".bin"), "w") do io
        write(io, Int64(ndims(A))); foreach(d -> write(io, Int64(d)), size(A))
        write(io, A isa AbstractArray{<:Complex} ? reinterpret(Float64, vec(collect(ComplexF64.(A)))) : vec(collect(Float64.(A))))
    end
end

# C2: segment table (pos, dir, n, t, normal per ray) of 257 rays of the Fibonacci disc + spot diagram
golden = π * (3 - sqrt(5)); n = 257
dl = SphericalDoubletLens(87.9e-3, -105.6e-3, Inf, 6e-3, 3e-3, 25.4e-3, 1.6456, 1.7168)
sd = Spotdetector(5e-3); translate3d!(sd, [0, BMO.thickness(dl) + 0.14368, 0])
sys = System([dl, sd])
seg = fill(NaN, 11, 8, n)
for i in 1:n
    r = sqrt((i - 0.5) / n) * 10e-3
    b = Beam(Ray([r * cos(i * golden), -0.05, r * sin(i * golden)], [0.0, 1.0, 0.0], 707e-9))
    solve_system!(sys, b)
    for (k, ray) in enumerate(BMO.rays(b))
        it = BMO.intersection(ray)
        seg[:, k, i] .= vcat(BMO.position(ray), BMO.direction(ray), BMO.refractive_index(ray),
            isnothing(it) ? [Inf, 0.0, 0.0, 0.0] : vcat(length(it), BMO.normal3d(it)))
    end
end
dump("c2_segments", seg)
dump("c2_spots", reduce(hcat, [collect(p) for p in sd.data]))

# C1: the Michelson interferometer of src/Workloads/michelson_wl.jl:8-67 (1" NBK7 cube splitter at 632.8 nm, right-angle prism
# mirror, two round plano mirrors, Photodetector 8 mm / 200 px) with a deterministic beamlet basis (support = [1, 0, 0])
cm = 1e-2; inch = 25.4e-3
NBK7 = DiscreteRefractiveIndex([632.8e-9], [1.51509])
rpm = RightAnglePrismMirror(25e-3, 25e-3); zrotate3d!(rpm, deg2rad(45)); translate3d!(rpm, [0, 33.5cm, 0])
cbs = CubeBeamsplitter(inch, NBK7); zrotate3d!(cbs, deg2rad(-90))
m1 = RoundPlanoMirror(inch, 5e-3); zrotate3d!(m1, deg2rad(-90)); translate3d!(m1, [22cm, 0, 0])
m2 = RoundPlanoMirror(inch, 5e-3); zrotate3d!(m2, deg2rad(-90)); translate3d!(m2, [12cm, 0, 0])
pd = Photodetector(8e-3, 200); translate3d!(pd, [0, -12cm, 0])
g_mirror, g_split, g_arm1, g_arm2, g_pd = ObjectGroup([rpm]), ObjectGroup([cbs]), ObjectGroup([m1]), ObjectGroup([m2]), ObjectGroup([pd])
michelson = System([g_mirror, g_split, g_arm1, g_arm2, g_pd])
origin = [18.81cm, 23.5cm, 0]
translate_to3d!(g_mirror, [0, -10cm, 0]); translate_to3d!(g_split, origin)
translate_to3d!(g_arm1, origin); translate_to3d!(g_arm2, origin)
translate3d!(g_arm1, [3.81cm / 2, 0, 0]); translate3d!(g_arm2, [0, 3.81cm / 2, 0]); zrotate3d!(g_arm2, deg2rad(90))
translate_to3d!(g_pd, origin); translate3d!(g_pd, [0, -3.81cm / 2, 0])
beam = GaussianBeamlet([0.0, 0, 0], [0.0, 1, 0], 632.8e-9, 5e-4; M2 = 2, support = [1.0, 0.0, 0.0])
solve_system!(michelson, beam)
dump("c1_field", pd.field)
dump("c1_power", [optical_power(pd)])
empty!(pd); translate3d!(m1, [5e-9, 0, 0]); solve_system!(michelson, beam)
dump("c1_field_shifted", pd.field)
dump("c1_power_shifted", [optical_power(pd)])
println("fixtures written to ", outdir)
