# GB25CUDA.jl — the ccall binding of libgb25cuda for the GordonBell25 drop-in (INTEGRATION.md section 2).
#
# UN-RUN: Julia is not installed in the image this repository was built in.  The same C ABI (include/gb25cuda.h) is
# exercised end to end by the Python ctypes binding (gb-25_b200/lib.py), which tests/ and bench.py drive.
# Usage:  ENV["LIBGB25CUDA"] = "/path/to/gb-25_b200/csrc/libgb25cuda.so"          (Float32 models)
#         ENV["LIBGB25CUDA_F64"] = "/path/to/gb-25_b200/csrc/libgb25cuda_f64.so"  (Float64 models, --float-type Float64)
#         include("julia/GB25CUDA.jl")
module GB25CUDA
using Oceananigans
using Oceananigans.Grids: halo_size, topology
import GordonBell25: first_time_step!, time_step!, loop!

const lib32 = get(ENV, "LIBGB25CUDA", "libgb25cuda")
const lib64 = get(ENV, "LIBGB25CUDA_F64", "libgb25cuda_f64")

struct Config   # layout of gb25_config (include/gb25cuda.h); model parameters are Float32 in both builds
    Nx::Cint; Ny::Cint; Nz::Cint; Hx::Cint; Hy::Cint; Hz::Cint
    topo_y::Cint; immersed::Cint; nsubsteps::Cint
    coriolis_scheme::Cint; fold_variant::Cint; south_inactive::Cint; cond_diff::Cint; eos_r0::Cint
    g::Cfloat; rho0::Cfloat; chi::Cfloat; dtau_frac::Cfloat; weno_eps::Cfloat
    Rx::Cint; Ry::Cint; rx::Cint; ry::Cint; device::Cint
    closure::Cint; kappa::Cfloat; nu::Cfloat     # 0 nothing | 1 VerticalScalarDiffusivity explicit | 2 vertically implicit
end
struct GridPtrs  # layout of gb25_grid: 18 pointers to gb25_real arrays + the Float32 averaging weights
    p::NTuple{19, Ptr{Cvoid}}
end

# FT = eltype(grid): Float32 -> libgb25cuda.so (all kernel generations), Float64 -> libgb25cuda_f64.so
mutable struct GB25CUDAModel{FT, M}
    cpu_model::M          # the unchanged Oceananigans model on CPU(): constructors stay as they are
    handle::Ptr{Cvoid}
end
libof(::GB25CUDAModel{Float32}) = lib32
libof(::GB25CUDAModel{Float64}) = lib64
libof(::Type{Float32}) = lib32
libof(::Type{Float64}) = lib64

# (Julia needs the library name of a ccall as a constant: one method per build)
macro gb25(FT, fn, ret, argt, args...)
    quote
        $(esc(FT)) === Float32 ? ccall(($(QuoteNode(fn)), lib32), $(esc(ret)), $(esc(argt)), $(map(esc, args)...)) :
                                 ccall(($(QuoteNode(fn)), lib64), $(esc(ret)), $(esc(argt)), $(map(esc, args)...))
    end
end
last_error(FT, h) = unsafe_string(@gb25 FT gb25_last_error Cstring (Ptr{Cvoid},) h)
check(FT, h, rc) = rc == 0 ? nothing : error("libgb25cuda: ", last_error(FT, h))
check(m::GB25CUDAModel{FT}, rc) where FT = check(FT, m.handle, rc)

# 2-D metric as a dense (Nx+2Hx, Ny+2Hy+1) array of FT (vectors of a LatitudeLongitudeGrid are broadcast)
function metric2d(FT, grid, a)
    Nx, Ny, _ = size(grid); Hx, Hy, _ = halo_size(grid)
    out = zeros(FT, Nx + 2Hx, Ny + 2Hy + 1)
    p = parent(a)
    if ndims(p) == 1
        n = min(length(p), size(out, 2)); out[:, 1:n] .= reshape(FT.(p[1:n]), 1, n)
    else
        n = min(size(p, 2), size(out, 2)); out[:, 1:n] .= FT.(p[:, 1:n])
    end
    return out
end

closure_code(::Nothing) = (0, 0f0, 0f0)
closure_code(c::VerticalScalarDiffusivity{<:Oceananigans.TurbulenceClosures.VerticallyImplicitTimeDiscretization}) = (2, Float32(c.κ.T), Float32(c.ν))
closure_code(c::VerticalScalarDiffusivity) = (1, Float32(c.κ.T), Float32(c.ν))

"""
    GB25CUDAModel(cpu_model; device = -1, partition = (1, 1, 0, 0))

`cpu_model` is the tile's model on `CPU()` (for a partitioned run: built on the tile's local grid, e.g. the local part of a
`Distributed(CPU(); partition = Partition(Rx, Ry, 1))` model); `partition = (Rx, Ry, rx, ry)` places it in the
`(Rx, Ry) = factors(Ndev)` decomposition of src/sharding_utils.jl:39-62, rank = rx + Rx*ry.
"""
function GB25CUDAModel(cpu_model; device = -1, partition = (1, 1, 0, 0))
    grid = cpu_model.grid
    FT = eltype(grid)
    ug = grid isa ImmersedBoundaryGrid ? grid.underlying_grid : grid
    Nx, Ny, Nz = size(grid); Hx, Hy, Hz = halo_size(grid)
    fs = cpu_model.free_surface
    w = Float32.(collect(fs.substepping.averaging_weights))
    fold = topology(grid, 2) != Bounded
    arrays = Any[metric2d(FT, ug, getproperty(ug, s)) for s in
                 (:Δxᶜᶜᵃ, :Δxᶠᶜᵃ, :Δxᶜᶠᵃ, :Δxᶠᶠᵃ, :Δyᶜᶜᵃ, :Δyᶠᶜᵃ, :Δyᶜᶠᵃ, :Δyᶠᶠᵃ, :Azᶜᶜᵃ, :Azᶠᶜᵃ, :Azᶜᶠᵃ, :Azᶠᶠᵃ)]
    push!(arrays, FT[2 * cpu_model.coriolis.rotation_rate * sind(Oceananigans.Grids.φnode(i, j, 1, ug, Face(), Face(), Center()))
                     for i in 1-Hx:Nx+Hx, j in 1-Hy:Ny+Hy+1])
    z = ug.z
    pad(v) = FT.(vcat(parent(v), fill(parent(v)[end], Nz + 2Hz + 1 - length(parent(v)))))
    append!(arrays, [pad(z.cᵃᵃᶠ), pad(z.cᵃᵃᶜ), pad(z.Δᵃᵃᶜ), pad(z.Δᵃᵃᶠ)])
    bottom = grid isa ImmersedBoundaryGrid ? metric2d(FT, ug, grid.immersed_boundary.bottom_height) : nothing
    push!(arrays, bottom === nothing ? FT[] : bottom); push!(arrays, w)
    Rx, Ry, rx, ry = partition
    cfg = Config(Nx, Ny, Nz, Hx, Hy, Hz, fold, bottom !== nothing, length(w), 1, 0, 1, 1, 0,
                 cpu_model.buoyancy.formulation.gravitational_acceleration, 1020, cpu_model.timestepper.χ,
                 fs.substepping.fractional_step_size, 1f-8, Rx, Ry, rx, ry, device,
                 closure_code(cpu_model.closure)...)     # (0, 0f0, 0f0) for `closure = nothing`
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve arrays begin
        ptrs = GridPtrs(ntuple(i -> isempty(arrays[i]) ? C_NULL : Ptr{Cvoid}(pointer(arrays[i])), 19))
        rc = @gb25 FT gb25_create Cint (Ref{Config}, Ref{GridPtrs}, Ref{Ptr{Cvoid}}) cfg ptrs h
    end
    check(FT, C_NULL, rc)
    m = GB25CUDAModel{FT, typeof(cpu_model)}(cpu_model, h[])
    finalizer(x -> (@gb25 FT gb25_destroy Cint (Ptr{Cvoid},) x.handle), m)
    upload!(m)
    return m
end

# ---- one Julia process driving every GPU of the node (sharding/sharded_baroclinic_instability_simulation_run.jl:49,
#      single_gpu_per_process = false): models[r] is the tile of rank r-1, each created with its own `device`
function connect_local!(models::Vector{<:GB25CUDAModel{FT}}) where FT
    hs = Ptr{Cvoid}[m.handle for m in models]
    check(models[1], @gb25 FT gb25_exchange_connect_local Cint (Ptr{Ptr{Cvoid}}, Cint) hs length(hs))
end
function loop!(models::Vector{<:GB25CUDAModel{FT}}, Ninner) where FT
    hs = Ptr{Cvoid}[m.handle for m in models]
    check(models[1], @gb25 FT gb25_loop_all Cint (Ptr{Ptr{Cvoid}}, Cint, Cfloat, Cint) hs length(hs) 0f0 Ninner)
end
first_time_step!(models::Vector{<:GB25CUDAModel}) = foreach(first_time_step!, models)   # all tiles enqueued before any is synchronised
time_step!(models::Vector{<:GB25CUDAModel}) = foreach(time_step!, models)

const FIELD_IDS = (u = 0, v = 1, w = 2, T = 3, S = 4, η = 14)   # gb25_field

# sync_states!(device <- host) and back: parents are passed as they are (column-major, halos included), one batch, one sync
function upload!(m::GB25CUDAModel{FT}) where FT
    Ψ = Oceananigans.fields(m.cpu_model)
    arrs = [Array{FT}(parent(Ψ[name])) for name in keys(FIELD_IDS)]
    ids = Cint[id for id in values(FIELD_IDS)]
    GC.@preserve arrs begin
        ptrs = Ptr{Cvoid}[pointer(a) for a in arrs]
        check(m, @gb25 FT gb25_set_fields Cint (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Cint) m.handle length(ids) ids ptrs 0)
    end
    c = m.cpu_model.clock
    @gb25 FT gb25_set_clock Cint (Ptr{Cvoid}, Cdouble, Clong, Cfloat) m.handle c.time c.iteration c.last_Δt
end
# set!(model, u = ..., v = ...): interior-shaped arrays, halos untouched, followed by update_state! as in Oceananigans
function Oceananigans.set!(m::GB25CUDAModel{FT}; kw...) where FT
    for (name, val) in kw
        a = Array{FT}(val)
        check(m, @gb25 FT gb25_set_interior Cint (Ptr{Cvoid}, Cint, Ptr{Cvoid}) m.handle FIELD_IDS[name] a)
    end
    Oceananigans.TimeSteppers.update_state!(m)
end
# FluxBoundaryCondition(values) at the top (side = 1) or bottom (0) of u, v, T, S — a 2-D parent-shaped array, or nothing
function set_flux_boundary_condition!(m::GB25CUDAModel{FT}, name, side, values) where FT
    p = values === nothing ? C_NULL : Ptr{Cvoid}(pointer(Array{FT}(values)))
    check(m, @gb25 FT gb25_set_flux_boundary_condition Cint (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}) m.handle FIELD_IDS[name] side p)
end
# (gb25_field id, parent array) of everything compare_states looks at (src/correctness.jl:28-90): fields(model),
# Gⁿ and G⁻ of u, v, T, S, and the filtered barotropic state
function compared_parents(m)
    M = m.cpu_model; ts = M.timestepper; fs = M.free_surface
    Ψ = Oceananigans.fields(M)
    list = Any[(id, parent(Ψ[name])) for (name, id) in pairs(FIELD_IDS)]
    for (k, name) in enumerate((:u, :v, :T, :S))
        push!(list, (5 + k, parent(ts.Gⁿ[name])))      # GB25_GN_U = 6, ...
        push!(list, (9 + k, parent(ts.G⁻[name])))      # GB25_GM_U = 10, ...
    end
    push!(list, (17, parent(fs.filtered_state.η)), (18, parent(fs.filtered_state.U)), (19, parent(fs.filtered_state.V)))
    return list
end
function download!(m::GB25CUDAModel{FT}) where FT   # so that GordonBell25.compare_states(m.cpu_model, vmodel) runs unchanged
    list = compared_parents(m)
    arrs = [Array{FT}(undef, size(p)) for (_, p) in list]
    ids = Cint[id for (id, _) in list]
    GC.@preserve arrs begin
        ptrs = Ptr{Cvoid}[pointer(a) for a in arrs]
        check(m, @gb25 FT gb25_get_fields Cint (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Cint) m.handle length(ids) ids ptrs 0)
    end
    for ((_, p), a) in zip(list, arrs)
        copyto!(p, a)
    end
end

# the drop-in: same names, same argument meaning as src/timestepping_utils.jl:21-45
first_time_step!(m::GB25CUDAModel{FT}) where FT =
    check(m, @gb25 FT gb25_first_time_step Cint (Ptr{Cvoid}, Cfloat) m.handle Float32(m.cpu_model.clock.last_Δt))
time_step!(m::GB25CUDAModel{FT}) where FT =
    check(m, @gb25 FT gb25_time_step Cint (Ptr{Cvoid}, Cfloat) m.handle 0f0)      # 0 => clock.last_Δt
loop!(m::GB25CUDAModel{FT}, Ninner) where FT =
    check(m, @gb25 FT gb25_loop Cint (Ptr{Cvoid}, Cfloat, Cint) m.handle 0f0 Ninner)
Oceananigans.initialize!(m::GB25CUDAModel{FT}) where FT = check(m, @gb25 FT gb25_initialize Cint (Ptr{Cvoid},) m.handle)
Oceananigans.TimeSteppers.update_state!(m::GB25CUDAModel{FT}) where FT = check(m, @gb25 FT gb25_update_state Cint (Ptr{Cvoid},) m.handle)
end
