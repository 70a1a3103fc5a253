# GB25CUDA.jl — the ccall binding of libgb25cuda for the GordonBell25 drop-in (INTEGRATION.md section 2).
#
# UN-RUN: Julia is not installed in the image this repository was built in.  The same C ABI (include/gb25cuda.h) is
# exercised end to end by the Python ctypes binding (gb-25_b200/lib.py), which tests/ and bench.py drive.
# Usage:  ENV["LIBGB25CUDA"] = "/path/to/gb-25_b200/csrc/libgb25cuda.so";  include("julia/GB25CUDA.jl")
module GB25CUDA
using Oceananigans
using Oceananigans.Grids: halo_size, topology
import GordonBell25: first_time_step!, time_step!, loop!

const lib = get(ENV, "LIBGB25CUDA", "libgb25cuda")

struct Config   # layout of gb25_config (include/gb25cuda.h)
    Nx::Cint; Ny::Cint; Nz::Cint; Hx::Cint; Hy::Cint; Hz::Cint
    topo_y::Cint; immersed::Cint; nsubsteps::Cint
    coriolis_scheme::Cint; fold_variant::Cint; south_inactive::Cint; cond_diff::Cint; eos_r0::Cint
    g::Cfloat; rho0::Cfloat; chi::Cfloat; dtau_frac::Cfloat; weno_eps::Cfloat
    Rx::Cint; Ry::Cint; rx::Cint; ry::Cint; device::Cint
    closure::Cint; kappa::Cfloat; nu::Cfloat     # 0 nothing | 1 VerticalScalarDiffusivity explicit | 2 vertically implicit
end
struct GridPtrs  # layout of gb25_grid: 19 Ptr{Cfloat}
    p::NTuple{19, Ptr{Cfloat}}
end

mutable struct GB25CUDAModel{M}
    cpu_model::M          # the unchanged Oceananigans model on CPU(): constructors stay as they are
    handle::Ptr{Cvoid}
end

check(h, rc) = rc == 0 ? nothing :
    error("libgb25cuda: ", unsafe_string(ccall((:gb25_last_error, lib), Cstring, (Ptr{Cvoid},), h)))

# 2-D metric as a dense (Nx+2Hx, Ny+2Hy+1) Float32 array (vectors of a LatitudeLongitudeGrid are broadcast)
function metric2d(grid, a)
    Nx, Ny, _ = size(grid); Hx, Hy, _ = halo_size(grid)
    out = zeros(Float32, Nx + 2Hx, Ny + 2Hy + 1)
    p = parent(a)
    if ndims(p) == 1
        n = min(length(p), size(out, 2)); out[:, 1:n] .= reshape(Float32.(p[1:n]), 1, n)
    else
        n = min(size(p, 2), size(out, 2)); out[:, 1:n] .= Float32.(p[:, 1:n])
    end
    return out
end

function GB25CUDAModel(cpu_model; device = -1)
    grid = cpu_model.grid
    ug = grid isa ImmersedBoundaryGrid ? grid.underlying_grid : grid
    Nx, Ny, Nz = size(grid); Hx, Hy, Hz = halo_size(grid)
    fs = cpu_model.free_surface
    w = Float32.(collect(fs.substepping.averaging_weights))
    fold = topology(grid, 2) != Bounded
    arrays = Any[metric2d(ug, getproperty(ug, s)) for s in
                 (:Δxᶜᶜᵃ, :Δxᶠᶜᵃ, :Δxᶜᶠᵃ, :Δxᶠᶠᵃ, :Δyᶜᶜᵃ, :Δyᶠᶜᵃ, :Δyᶜᶠᵃ, :Δyᶠᶠᵃ, :Azᶜᶜᵃ, :Azᶠᶜᵃ, :Azᶜᶠᵃ, :Azᶠᶠᵃ)]
    push!(arrays, Float32[2 * cpu_model.coriolis.rotation_rate * sind(Oceananigans.Grids.φnode(i, j, 1, ug, Face(), Face(), Center()))
                          for i in 1-Hx:Nx+Hx, j in 1-Hy:Ny+Hy+1])
    z = ug.z
    pad(v) = Float32.(vcat(parent(v), fill(parent(v)[end], Nz + 2Hz + 1 - length(parent(v)))))
    append!(arrays, [pad(z.cᵃᵃᶠ), pad(z.cᵃᵃᶜ), pad(z.Δᵃᵃᶜ), pad(z.Δᵃᵃᶠ)])
    bottom = grid isa ImmersedBoundaryGrid ? metric2d(ug, grid.immersed_boundary.bottom_height) : nothing
    push!(arrays, bottom === nothing ? Float32[] : bottom); push!(arrays, w)
    cfg = Config(Nx, Ny, Nz, Hx, Hy, Hz, fold, bottom !== nothing, length(w), 1, 0, 1, 1, 0,
                 cpu_model.buoyancy.formulation.gravitational_acceleration, 1020, cpu_model.timestepper.χ,
                 fs.substepping.fractional_step_size, 1f-8, 1, 1, 0, 0, device,
                 closure_code(cpu_model.closure)...)     # (0, 0f0, 0f0) for `closure = nothing`
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve arrays begin
        ptrs = GridPtrs(ntuple(i -> isempty(arrays[i]) ? Ptr{Cfloat}(C_NULL) : pointer(arrays[i]), 19))
        rc = ccall((:gb25_create, lib), Cint, (Ref{Config}, Ref{GridPtrs}, Ref{Ptr{Cvoid}}), cfg, ptrs, h)
    end
    check(C_NULL, rc)
    m = GB25CUDAModel(cpu_model, h[])
    finalizer(x -> ccall((:gb25_destroy, lib), Cint, (Ptr{Cvoid},), x.handle), m)
    upload!(m)
    return m
end

const FIELD_IDS = (u = 0, v = 1, w = 2, T = 3, S = 4, η = 14)   # gb25_field
closure_code(::Nothing) = (0, 0f0, 0f0)
closure_code(c::VerticalScalarDiffusivity{<:Oceananigans.TurbulenceClosures.VerticallyImplicitTimeDiscretization}) = (2, Float32(c.κ.T), Float32(c.ν))
closure_code(c::VerticalScalarDiffusivity) = (1, Float32(c.κ.T), Float32(c.ν))

# sync_states!(device <- host) and back: parents are passed as they are (column-major, halos included)
function upload!(m)
    for (name, id) in pairs(FIELD_IDS)
        a = Array{Float32}(parent(Oceananigans.fields(m.cpu_model)[name]))
        check(m.handle, ccall((:gb25_set_field, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cfloat}), m.handle, id, a))
    end
    c = m.cpu_model.clock
    ccall((:gb25_set_clock, lib), Cint, (Ptr{Cvoid}, Cdouble, Clong, Cfloat), m.handle, c.time, c.iteration, c.last_Δt)
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
function download!(m)   # so that GordonBell25.compare_states(m.cpu_model, vmodel) runs unchanged
    for (id, p) in compared_parents(m)
        a = Array{Float32}(undef, size(p))
        check(m.handle, ccall((:gb25_get_field, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cfloat}), m.handle, id, a))
        copyto!(p, a)
    end
end

# the drop-in: same names, same argument meaning as src/timestepping_utils.jl:21-45
first_time_step!(m::GB25CUDAModel) =
    check(m.handle, ccall((:gb25_first_time_step, lib), Cint, (Ptr{Cvoid}, Cfloat), m.handle, m.cpu_model.clock.last_Δt))
time_step!(m::GB25CUDAModel) =
    check(m.handle, ccall((:gb25_time_step, lib), Cint, (Ptr{Cvoid}, Cfloat), m.handle, 0f0))      # 0 => clock.last_Δt
loop!(m::GB25CUDAModel, Ninner) =
    check(m.handle, ccall((:gb25_loop, lib), Cint, (Ptr{Cvoid}, Cfloat, Cint), m.handle, 0f0, Ninner))
Oceananigans.initialize!(m::GB25CUDAModel) = check(m.handle, ccall((:gb25_initialize, lib), Cint, (Ptr{Cvoid},), m.handle))
Oceananigans.TimeSteppers.update_state!(m::GB25CUDAModel) = check(m.handle, ccall((:gb25_update_state, lib), Cint, (Ptr{Cvoid},), m.handle))
end
