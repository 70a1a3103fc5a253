# correctness_libgb25.jl — the reference's correctness protocol with libgb25cuda in the `rmodel` seat.
#
# Mirrors /root/reference/correctness/correctness_baroclinic_instability_simulation_run.jl (stages and tolerances:
# lines 14-17, 33-43, 45-102): same model on both sides, random u, v, states synchronised, then
# initialize! / update_state! / first_time_step! / time_step! x10 / loop!, with GordonBell25.compare_states after every
# stage (halos included, rtol = sqrt(eps(Float32)), atol = 0).  UN-RUN here (no Julia in the build image); running it
# with Oceananigans =0.96.26 is how the fidelity decisions U1-U15 of DESIGN.md section 6 get confirmed or corrected.
#
#   LIBGB25CUDA=/path/to/libgb25cuda.so julia --project=/path/to/GB-25 julia/correctness_libgb25.jl [simple_lat_lon|gaussian_islands]
using GordonBell25
using GordonBell25: first_time_step!, time_step!, loop!
using Oceananigans
using Random

include(joinpath(@__DIR__, "GB25CUDA.jl"))

const FT = Float32
Oceananigans.defaults.FloatType = FT
throw_error = true
include_halos = true
rtol = sqrt(eps(FT))
atol = 0

grid_type = length(ARGS) >= 1 ? Symbol(ARGS[1]) : :simple_lat_lon
Nx, Ny, Nz = 128, 64, 8                     # BASELINE configs[0]
model_kw = (; halo = (8, 8, 8), Δt = 1e-9, grid_type)

vmodel = GordonBell25.baroclinic_instability_model(CPU(), Nx, Ny, Nz; model_kw...)
cmodel = GordonBell25.baroclinic_instability_model(CPU(), Nx, Ny, Nz; model_kw...)

Random.seed!(42)
ui = 1e-3 .* rand(FT, size(vmodel.velocities.u)...)
vi = 1e-3 .* rand(FT, size(vmodel.velocities.v)...)
set!(vmodel, u = ui, v = vi)
GordonBell25.sync_states!(cmodel, vmodel)
rmodel = GB25CUDA.GB25CUDAModel(cmodel)     # uploads grid products and state

check(label) = begin
    GB25CUDA.download!(rmodel)
    @info label
    GordonBell25.compare_states(rmodel.cpu_model, vmodel; include_halos, throw_error, rtol, atol)
end

check("After syncing and uploading:")
Oceananigans.initialize!(rmodel); Oceananigans.initialize!(vmodel)
check("After initialize!:")
Oceananigans.TimeSteppers.update_state!(rmodel); Oceananigans.TimeSteppers.update_state!(vmodel)
check("After update_state!:")
first_time_step!(rmodel); first_time_step!(vmodel)
check("After first time step:")
for n in 1:10
    time_step!(rmodel); time_step!(vmodel)
end
check("After 10 more steps:")
Nt = 100
loop!(rmodel, Nt); loop!(vmodel, Nt)
check("After a loop of $Nt steps:")
