# bench_cpu_oceananigans.jl — the CPU baseline the north-star asks for: Oceananigans CPU() stepping the same
# baroclinic-instability model through GordonBell25.loop!, timed on all host cores.
#
# UN-RUN here (no Julia in the build image; bench.py times the C++ oracle port instead and labels it "port").
# Run on the GPU box's host to fill in the true number:
#
#   JULIA_NUM_THREADS=$(nproc) julia --project=/path/to/GB-25 julia/bench_cpu_oceananigans.jl 360 150 50 25
#
# Prints one JSON line in the shape of bench.py's cpu_baseline object (kind = "reference").
using GordonBell25
using GordonBell25: first_time_step!, loop!
using Oceananigans
using Random

const FT = Float32
Oceananigans.defaults.FloatType = FT
Nx, Ny, Nz, Nt = length(ARGS) >= 4 ? parse.(Int, ARGS[1:4]) : (360, 150, 50, 25)   # a 1/16 tile of BASELINE configs[1]

model = GordonBell25.baroclinic_instability_model(CPU(), Nx, Ny, Nz; halo = (8, 8, 8), Δt = 60 * 4, grid_type = :gaussian_islands)
GordonBell25.set_baroclinic_instability!(model)
Random.seed!(42)
set!(model, u = 1e-3 .* rand(FT, size(model.velocities.u)...), v = 1e-3 .* rand(FT, size(model.velocities.v)...))

first_time_step!(model)
loop!(model, 2)                                   # warm-up (compilation)
t = @elapsed loop!(model, Nt)
value = Nx * Ny * Nz * Nt / t
println("{\"value\": $value, \"unit\": \"cell-steps/s\", \"cores\": $(Threads.nthreads()), \"kind\": \"reference\", ",
        "\"sample\": \"Oceananigans CPU() gaussian_islands $(Nx)x$(Ny)x$(Nz), $Nt steps in $(round(t, digits = 2)) s\"}")
