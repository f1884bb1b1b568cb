# PNB200.jl -- the Julia glue that makes libpnb200.so a drop-in for the hot path of
# PointNeighbors.jl (v0.6.7) on a new device array type.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia` binary (SURVEY.md 8c).
# Every method below is a 1:1 mirror of a function of the Python host layer
# (pointneighbors.jl_b200/pnb200/api.py), which IS tested against the CPU oracle on CPU and on
# B200 -- the arithmetic, the error texts and the call order live in the C library, not here.
#
# Dispatch axis (same as the reference): the array type of the coordinates.
#   default_backend(x::AbstractGPUArray) picks the backend        (src/util.jl:81-83)
#   Adapt.adapt(backend, nhs) moves a search to the device          (src/gpu.jl:10-35)
# so a user changes `CuArray`/`CUDABackend()` to `B200Array`/`B200Backend()` and nothing else.
module PNB200

using PointNeighbors
using PointNeighbors: GridNeighborhoodSearch, FullGridCellList, SpatialHashingCellList, PeriodicBox,
                      PrecomputedNeighborhoodSearch, AbstractNeighborhoodSearch
using PointNeighbors.Adapt
using GPUArraysCore: AbstractGPUArray

import PointNeighbors: initialize!, update!, foreach_point_neighbor, copy_neighborhood_search,
                       freeze_neighborhood_search, requires_update, search_radius, default_backend

export B200Array, B200Backend, CountNeighbors, NBodyGravity, WCSPHInteract,
       TLSPHDeformationGradient, TLSPHInteract, TlsphParams, compute_pk1_corrected!,
       compute_pressure!, set_exact_arithmetic

const libpnb200 = get(ENV, "PNB200_LIB",
                      joinpath(@__DIR__, "..", "pnb200", "libpnb200.so"))

# ---------------------------------------------------------------------------------------------
# status -> the exception the reference raises (include/pnb200.h, pnb_status)
# ---------------------------------------------------------------------------------------------
last_error() = unsafe_string(ccall((:pnb_last_error, libpnb200), Cstring, ()))

function check(status::Integer)
    status == 0 && return nothing
    msg = last_error()
    status == 2 && throw(ArgumentError(msg))                 # PNB_ERR_ARG
    status == 4 && throw(BoundsError(msg))                   # PNB_ERR_BOUNDS
    error(msg)                                               # DOMAIN / LIST_FULL / STATE / CUDA
end

# ---------------------------------------------------------------------------------------------
# device array: owns a pnb_malloc'ed buffer (there is no CUDA.jl in the loop)
# ---------------------------------------------------------------------------------------------
mutable struct B200Array{T, N} <: AbstractGPUArray{T, N}
    ptr  :: Ptr{Cvoid}
    dims :: NTuple{N, Int}

    function B200Array{T, N}(::UndefInitializer, dims::NTuple{N, Int}) where {T, N}
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:pnb_malloc, libpnb200), Cint, (Ref{Ptr{Cvoid}}, Int64),
                    ref, prod(dims) * sizeof(T)))
        a = new{T, N}(ref[], dims)
        finalizer(x -> ccall((:pnb_free, libpnb200), Cint, (Ptr{Cvoid},), x.ptr), a)
        return a
    end
end

B200Array{T}(::UndefInitializer, dims::Int...) where {T} = B200Array{T, length(dims)}(undef, dims)
Base.size(a::B200Array) = a.dims
Base.similar(a::B200Array{T}, ::Type{S}, dims::Dims) where {T, S} = B200Array{S, length(dims)}(undef, dims)
Base.pointer(a::B200Array) = a.ptr

function B200Array(h::Array{T, N}) where {T, N}            # host -> device
    a = B200Array{T, N}(undef, size(h))
    check(ccall((:pnb_memcpy_h2d, libpnb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
                a.ptr, h, sizeof(h), C_NULL))
    return a
end

function Base.Array(a::B200Array{T, N}) where {T, N}       # device -> host
    h = Array{T, N}(undef, a.dims)
    check(ccall((:pnb_memcpy_d2h, libpnb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
                h, a.ptr, sizeof(h), C_NULL))
    return h
end

function Base.fill!(a::B200Array{T}, v) where {T}
    iszero(v) || error("B200Array only supports fill!(a, 0)")
    check(ccall((:pnb_memset, libpnb200), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}),
                a.ptr, 0, prod(a.dims) * sizeof(T), C_NULL))
    return a
end

# the backend tag; subtyping AbstractThreadingBackend makes it a legal `parallelization_backend`
# keyword (src/util.jl:36,70)
struct B200Backend <: PointNeighbors.AbstractThreadingBackend end
default_backend(::B200Array) = B200Backend()
Adapt.adapt_storage(::B200Backend, a::Array) = B200Array(a)
Adapt.adapt_storage(::Type{<:B200Array}, a::Array) = B200Array(a)
Adapt.adapt_storage(::Type{Array}, a::B200Array) = Array(a)

# ---------------------------------------------------------------------------------------------
# GridNeighborhoodSearch + FullGridCellList on the device
# ---------------------------------------------------------------------------------------------
# adapt(backend, nhs) returns this type: the reference's host scalars + the device handle.
mutable struct B200GridNeighborhoodSearch{NDIMS, ELTYPE, PB, US} <: AbstractNeighborhoodSearch
    handle          :: Ptr{Cvoid}
    search_radius   :: ELTYPE
    periodic_box    :: PB
    n_cells         :: NTuple{NDIMS, Int}
    cell_size       :: NTuple{NDIMS, ELTYPE}
    update_strategy :: US
    host            :: Any        # the GridNeighborhoodSearch it was adapted from (for copy_...)
    y_ref           :: Any        # the coordinates of the last initialize!/update!: kept alive,
                                  # the library checks that sweeps are called with this array
end

Base.ndims(::B200GridNeighborhoodSearch{NDIMS}) where {NDIMS} = NDIMS
requires_update(::B200GridNeighborhoodSearch) = (false, true)          # src/nhs_grid.jl:133

function Adapt.adapt_structure(::B200Backend, nhs::GridNeighborhoodSearch{NDIMS}) where {NDIMS}
    T = typeof(nhs.search_radius)
    T in (Float32, Float64) ||
        throw(ArgumentError("the B200 path computes in Float32 or Float64: pass such a `search_radius`"))
    cl = nhs.cell_list
    if cl isa SpatialHashingCellList
        # src/cell_lists/spatial_hashing.jl, src/gpu.jl:37-44: the table lives in the library
        T === Float32 ||
            throw(ArgumentError("the hashed cell list is available for Float32 searches"))
        box = nhs.periodic_box
        bmin = isnothing(box) ? C_NULL : collect(T, box.min_corner)
        bmax = isnothing(box) ? C_NULL : collect(T, box.max_corner)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:pnb_grid_create_hashed_f32, libpnb200), Cint,
                    (Cint, Cfloat, Int64, Ptr{Cfloat}, Ptr{Cfloat}, Ref{Ptr{Cvoid}}),
                    NDIMS, nhs.search_radius, cl.list_size, bmin, bmax, ref))
        out = B200GridNeighborhoodSearch{NDIMS, T, typeof(box), typeof(nhs.update_strategy)}(
            ref[], nhs.search_radius, box, nhs.n_cells, nhs.cell_size, nhs.update_strategy, nhs, nothing)
        finalizer(x -> ccall((:pnb_grid_destroy, libpnb200), Cvoid, (Ptr{Cvoid},), x.handle), out)
        return out
    end
    cl isa FullGridCellList ||
        throw(ArgumentError("only the FullGridCellList and the SpatialHashingCellList are GPU-compatible (src/cell_lists/dictionary.jl:8-10)"))
    if T === Float32 && eltype(cl.min_corner) === Float64
        # mixed precision (docs/literate/src/tut_gpu_usage.jl:45-50): Float64 corners and
        # coordinates, Float32 radius; used with B200Array{Float64} coordinates afterwards
        box = nhs.periodic_box
        bmin = isnothing(box) ? C_NULL : collect(Float32, box.min_corner)
        bmax = isnothing(box) ? C_NULL : collect(Float32, box.max_corner)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:pnb_grid_create_padded_mixed, libpnb200), Cint,
                    (Cint, Cfloat, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cfloat}, Ptr{Cfloat}, Ref{Ptr{Cvoid}}),
                    NDIMS, nhs.search_radius, collect(Float64, cl.min_corner),
                    collect(Float64, cl.max_corner), bmin, bmax, ref))
        out = B200GridNeighborhoodSearch{NDIMS, Float64, typeof(box), typeof(nhs.update_strategy)}(
            ref[], nhs.search_radius, box, nhs.n_cells, nhs.cell_size, nhs.update_strategy, nhs, nothing)
        finalizer(x -> ccall((:pnb_grid_destroy, libpnb200), Cvoid, (Ptr{Cvoid},), x.handle), out)
        return out
    end
    eltype(cl.min_corner) == T ||
        throw(ArgumentError("cell list corners and `search_radius` must have the same element type"))
    r = nhs.search_radius
    # cell_list.min_corner / max_corner are the PADDED corners (src/cell_lists/full_grid.jl:66-67):
    # hand them over as they are, the library continues at :74 (grid size from the stored corners)
    min_corner = collect(T, cl.min_corner)
    max_corner = collect(T, cl.max_corner)
    box = nhs.periodic_box
    bmin = isnothing(box) ? C_NULL : collect(T, box.min_corner)
    bmax = isnothing(box) ? C_NULL : collect(T, box.max_corner)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    if T === Float32
        check(ccall((:pnb_grid_create_padded_f32, libpnb200), Cint,
                    (Cint, Cfloat, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ref{Ptr{Cvoid}}),
                    NDIMS, r, min_corner, max_corner, bmin, bmax, ref))
    else
        check(ccall((:pnb_grid_create_padded_f64, libpnb200), Cint,
                    (Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Ptr{Cvoid}}),
                    NDIMS, r, min_corner, max_corner, bmin, bmax, ref))
    end
    out = B200GridNeighborhoodSearch{NDIMS, T, typeof(box), typeof(nhs.update_strategy)}(
        ref[], r, box, nhs.n_cells, nhs.cell_size, nhs.update_strategy, nhs, nothing)
    finalizer(x -> ccall((:pnb_grid_destroy, libpnb200), Cvoid, (Ptr{Cvoid},), x.handle), out)
    return out
end

index_vector(::Nothing) = (C_NULL, 0)
function index_vector(idx)                                  # eachindex_y / points -> device Int32
    v = B200Array(Int32.(collect(idx)))
    return (v, length(idx))
end
is_all(idx, n) = idx isa Base.OneTo ? length(idx) == n : (idx == 1:n)

# initialize!(nhs, x, y; eachindex_y)      src/nhs_grid.jl:220-225, 255-281
function initialize!(nhs::B200GridNeighborhoodSearch{NDIMS, T}, x::B200Array{T, 2},
                     y::B200Array{T, 2}; parallelization_backend = default_backend(x),
                     eachindex_y = axes(y, 2)) where {NDIMS, T <: Union{Float32, Float64}}
    n = size(y, 2)
    iv, ni = is_all(eachindex_y, n) ? (C_NULL, 0) : index_vector(eachindex_y)
    ivp = iv === C_NULL ? C_NULL : iv.ptr
    # the element type of the search (radius, cell list, coordinates) selects the entry point
    GC.@preserve iv if T === Float32
        check(ccall((:pnb_grid_build_f32, libpnb200), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}),
                    nhs.handle, y.ptr, n, ivp, ni, 1, C_NULL))
    else
        check(ccall((:pnb_grid_build_f64, libpnb200), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}),
                    nhs.handle, y.ptr, n, ivp, ni, 1, C_NULL))
    end
    nhs.y_ref = y
    return nhs
end

# update!(nhs, x, y; points_moving, eachindex_y)      src/nhs_grid.jl:283-292
# blocking = false (no reference counterpart, the reference synchronises after every launch,
# src/util.jl:166-170): the rebuild is only enqueued; a domain error is raised by the next
# blocking call on the search or by check!(nhs).  Float32 full rebuilds only.
function update!(nhs::B200GridNeighborhoodSearch{NDIMS, T}, x::B200Array{T, 2},
                 y::B200Array{T, 2}; points_moving = (true, true),
                 parallelization_backend = default_backend(x),
                 eachindex_y = axes(y, 2), blocking = true) where {NDIMS, T <: Union{Float32, Float64}}
    points_moving[2] || return nhs
    if !blocking && T === Float32 && is_all(eachindex_y, size(y, 2))
        check(ccall((:pnb_grid_build_async_f32, libpnb200), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
                    nhs.handle, y.ptr, size(y, 2), C_NULL))
        nhs.y_ref = y
        return nhs
    end
    return initialize!(nhs, x, y; eachindex_y)
end

# initialize_grid! / update_grid!(nhs, y; eachindex_y)      src/nhs_grid.jl:227-281, 470-477
# (the cell-list part of initialize! / update!; ParallelUpdate's update_grid! is a full rebuild)
initialize_grid!(nhs::B200GridNeighborhoodSearch, y::B200Array; parallelization_backend = default_backend(y),
                 eachindex_y = axes(y, 2)) = initialize!(nhs, y, y; eachindex_y)
update_grid!(nhs::B200GridNeighborhoodSearch, y::B200Array; parallelization_backend = default_backend(y),
             eachindex_y = axes(y, 2)) = initialize!(nhs, y, y; eachindex_y)

# foreach_point_neighbor_unsafe      src/neighborhood_search.jl:204-234
# (the reference skips the bounds checks inside its kernel; the device kernels here never index
#  out of bounds on an initialized search, so both names run the same code)
foreach_point_neighbor_unsafe(f, x::B200Array, y::B200Array, nhs; kwargs...) =
    foreach_point_neighbor(f, x, y, nhs; kwargs...)

# synchronise and raise what a blocking update! would have raised
function check!(nhs::B200GridNeighborhoodSearch)
    check(ccall((:pnb_grid_check, libpnb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), nhs.handle, C_NULL))
    return nhs
end

# copy_neighborhood_search(nhs, search_radius, n_points)      src/nhs_grid.jl:640-649
function copy_neighborhood_search(nhs::B200GridNeighborhoodSearch, search_radius, n_points;
                                  eachpoint = 1:n_points)
    host = copy_neighborhood_search(nhs.host, search_radius, n_points; eachpoint)
    return Adapt.adapt_structure(B200Backend(), host)
end

# ---------------------------------------------------------------------------------------------
# closures with a fused kernel (named functors: anonymous Julia closures cannot be recognised)
# ---------------------------------------------------------------------------------------------
struct CountNeighbors{A}                    # benchmarks/count_neighbors.jl:24-27
    n_neighbors :: A                        # B200Array{Int64, 1}
end
struct NBodyGravity{A, M}                   # benchmarks/n_body.jl:38-48
    dv   :: A
    mass :: M
    G    :: Float32
end
struct WcsphParams                          # pnb_wcsph_params
    smoothing_length :: Cfloat
    sound_speed      :: Cfloat
    alpha            :: Cfloat
    beta             :: Cfloat
    epsilon          :: Cfloat
    delta            :: Cfloat
    kernel_norm      :: Cfloat
end
struct WCSPHInteract{A}                     # benchmarks/smoothed_particle_hydrodynamics.jl:45-102
    dv :: A; v_x :: A; v_y :: A
    mass_x :: Any; mass_y :: Any; pressure_x :: Any; pressure_y :: Any
    params :: WcsphParams
end

set_exact_arithmetic(on::Bool) = ccall((:pnb_set_exact_arithmetic, libpnb200), Cvoid, (Cint,), on)

points_arg(points, nx) = is_all(points, nx) ? (C_NULL, 0) : index_vector(points)

# foreach_point_neighbor(f, x, y, nhs; points)      src/neighborhood_search.jl:183-201
function foreach_point_neighbor(f::CountNeighbors, x::B200Array{Float32, 2},
                                y::B200Array{Float32, 2}, nhs::B200GridNeighborhoodSearch;
                                parallelization_backend = default_backend(x),
                                points = axes(x, 2))
    pv, np = points_arg(points, size(x, 2))
    GC.@preserve pv check(ccall((:pnb_count_neighbors_f32, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint,
                 Ptr{Cvoid}, Ptr{Cvoid}),
                nhs.handle, x.ptr, size(x, 2), y.ptr, size(y, 2),
                pv === C_NULL ? C_NULL : pv.ptr, np, 1, f.n_neighbors.ptr, C_NULL))
    return nothing
end

# Float64 searches: neighbour counts (and neighbour lists, below) exist in Float64; the fused
# n-body / WCSPH / TLSPH closures are Float32 only.
function foreach_point_neighbor(f::CountNeighbors, x::B200Array{Float64, 2},
                                y::B200Array{Float64, 2},
                                nhs::B200GridNeighborhoodSearch{NDIMS, Float64};
                                parallelization_backend = default_backend(x),
                                points = axes(x, 2)) where {NDIMS}
    pv, np = points_arg(points, size(x, 2))
    GC.@preserve pv check(ccall((:pnb_count_neighbors_f64, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint,
                 Ptr{Cvoid}, Ptr{Cvoid}),
                nhs.handle, x.ptr, size(x, 2), y.ptr, size(y, 2),
                pv === C_NULL ? C_NULL : pv.ptr, np, 1, f.n_neighbors.ptr, C_NULL))
    return nothing
end

function foreach_point_neighbor(f::NBodyGravity, x::B200Array{Float32, 2},
                                y::B200Array{Float32, 2}, nhs::B200GridNeighborhoodSearch;
                                parallelization_backend = default_backend(x),
                                points = axes(x, 2))
    pv, np = points_arg(points, size(x, 2))
    GC.@preserve pv check(ccall((:pnb_nbody_f32, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint,
                 Ptr{Cvoid}, Cfloat, Ptr{Cvoid}, Ptr{Cvoid}),
                nhs.handle, x.ptr, size(x, 2), y.ptr, size(y, 2),
                pv === C_NULL ? C_NULL : pv.ptr, np, 1, f.mass.ptr, f.G, f.dv.ptr, C_NULL))
    return nothing
end

function foreach_point_neighbor(f::WCSPHInteract, x::B200Array{Float32, 2},
                                y::B200Array{Float32, 2}, nhs::B200GridNeighborhoodSearch;
                                parallelization_backend = default_backend(x),
                                points = axes(x, 2))
    pv, np = points_arg(points, size(x, 2))
    prm = Ref(f.params)
    GC.@preserve pv check(ccall((:pnb_wcsph_interact_f32, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint,
                 Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                 Ref{WcsphParams}, Ptr{Cvoid}, Ptr{Cvoid}),
                nhs.handle, x.ptr, size(x, 2), y.ptr, size(y, 2),
                pv === C_NULL ? C_NULL : pv.ptr, np, 1, f.v_x.ptr, f.v_y.ptr, f.mass_x.ptr,
                f.mass_y.ptr, f.pressure_x.ptr, f.pressure_y.ptr, prm, f.dv.ptr, C_NULL))
    return nothing
end

# stream-ordered form (x === y, all points): nothing is synchronised, check!(nhs) settles it
function foreach_point_neighbor_async(f::WCSPHInteract, y::B200Array{Float32, 2},
                                      nhs::B200GridNeighborhoodSearch)
    prm = Ref(f.params)
    check(ccall((:pnb_wcsph_interact_async_f32, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                 Ref{WcsphParams}, Ptr{Cvoid}, Ptr{Cvoid}),
                nhs.handle, y.ptr, size(y, 2), f.v_y.ptr, f.mass_y.ptr, f.pressure_y.ptr, prm,
                f.dv.ptr, C_NULL))
    return nothing
end

# Float64 searches: the fused closures with Float64 arrays (bit-identical to the Float64 oracle)
struct NBodyGravity64{A, M}
    dv   :: A
    mass :: M
    G    :: Float64
end
struct WcsphParams64                        # pnb_wcsph_params_f64
    smoothing_length :: Cdouble
    sound_speed      :: Cdouble
    alpha            :: Cdouble
    beta             :: Cdouble
    epsilon          :: Cdouble
    delta            :: Cdouble
    kernel_norm      :: Cdouble
end
struct WCSPHInteract64{A}
    dv :: A; v_x :: A; v_y :: A
    mass_x :: Any; mass_y :: Any; pressure_x :: Any; pressure_y :: Any
    params :: WcsphParams64
end
function foreach_point_neighbor(f::NBodyGravity64, x::B200Array{Float64, 2}, y::B200Array{Float64, 2},
                                nhs::B200GridNeighborhoodSearch{NDIMS, Float64};
                                parallelization_backend = default_backend(x),
                                points = axes(x, 2)) where {NDIMS}
    pv, np = points_arg(points, size(x, 2))
    GC.@preserve pv check(ccall((:pnb_nbody_f64, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint,
                 Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}),
                nhs.handle, x.ptr, size(x, 2), y.ptr, size(y, 2),
                pv === C_NULL ? C_NULL : pv.ptr, np, 1, f.mass.ptr, f.G, f.dv.ptr, C_NULL))
    return nothing
end
function foreach_point_neighbor(f::WCSPHInteract64, x::B200Array{Float64, 2}, y::B200Array{Float64, 2},
                                nhs::B200GridNeighborhoodSearch{NDIMS, Float64};
                                parallelization_backend = default_backend(x),
                                points = axes(x, 2)) where {NDIMS}
    pv, np = points_arg(points, size(x, 2))
    prm = Ref(f.params)
    GC.@preserve pv check(ccall((:pnb_wcsph_interact_f64, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint,
                 Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                 Ref{WcsphParams64}, Ptr{Cvoid}, Ptr{Cvoid}),
                nhs.handle, x.ptr, size(x, 2), y.ptr, size(y, 2),
                pv === C_NULL ? C_NULL : pv.ptr, np, 1, f.v_x.ptr, f.v_y.ptr, f.mass_x.ptr,
                f.mass_y.ptr, f.pressure_x.ptr, f.pressure_y.ptr, prm, f.dv.ptr, C_NULL))
    return nothing
end

# Mixed precision (Float64 coordinates, Float32 search radius, tut_gpu_usage.jl:45-50): the closure
# receives Float32 pos_diff / distance, so the Float32 closures apply to Float64 coordinates; the
# library rejects a handle that is not a mixed-precision search.
function foreach_point_neighbor(f::NBodyGravity, x::B200Array{Float64, 2}, y::B200Array{Float64, 2},
                                nhs::B200GridNeighborhoodSearch;
                                parallelization_backend = default_backend(x),
                                points = axes(x, 2))
    pv, np = points_arg(points, size(x, 2))
    GC.@preserve pv check(ccall((:pnb_nbody_mixed, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint,
                 Ptr{Cvoid}, Cfloat, Ptr{Cvoid}, Ptr{Cvoid}),
                nhs.handle, x.ptr, size(x, 2), y.ptr, size(y, 2),
                pv === C_NULL ? C_NULL : pv.ptr, np, 1, f.mass.ptr, f.G, f.dv.ptr, C_NULL))
    return nothing
end
function foreach_point_neighbor(f::WCSPHInteract, x::B200Array{Float64, 2}, y::B200Array{Float64, 2},
                                nhs::B200GridNeighborhoodSearch;
                                parallelization_backend = default_backend(x),
                                points = axes(x, 2))
    pv, np = points_arg(points, size(x, 2))
    prm = Ref(f.params)
    GC.@preserve pv check(ccall((:pnb_wcsph_interact_mixed, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint,
                 Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                 Ref{WcsphParams}, Ptr{Cvoid}, Ptr{Cvoid}),
                nhs.handle, x.ptr, size(x, 2), y.ptr, size(y, 2),
                pv === C_NULL ? C_NULL : pv.ptr, np, 1, f.v_x.ptr, f.v_y.ptr, f.mass_x.ptr,
                f.mass_y.ptr, f.pressure_x.ptr, f.pressure_y.ptr, prm, f.dv.ptr, C_NULL))
    return nothing
end

# The WCSPH step from HOST arrays, pipelined inside the library (pnb_hoststep_*): what a host-side
# caller does per step -- copy coordinates and state to the device, update!, interact!, copy dv
# back -- with the copies of neighbouring steps overlapping the kernels.  Host arrays should be
# pinned (pnb_malloc_host) for the overlap.
mutable struct HostStepper
    handle :: Ptr{Cvoid}
    nhs    :: B200GridNeighborhoodSearch
    keep   :: Vector{Any}
    function HostStepper(nhs::B200GridNeighborhoodSearch, n_points::Integer)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:pnb_hoststep_create, libpnb200), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}),
                    nhs.handle, n_points, ref))
        s = new(ref[], nhs, Any[])
        finalizer(z -> ccall((:pnb_hoststep_destroy, libpnb200), Cvoid, (Ptr{Cvoid},), z.handle), s)
        return s
    end
end
# pressure === nothing: computed on the device from the density row of v (compute_pressure!);
# needs set_state_equation! first
function submit!(s::HostStepper, y::Matrix{Float32}, v::Matrix{Float32},
                 pressure::Union{Nothing, Vector{Float32}}, dv::Matrix{Float32}, params::WcsphParams;
                 mass::Union{Nothing, Vector{Float32}} = nothing)
    push!(s.keep, (y, v, pressure, dv, mass)); length(s.keep) > 3 && popfirst!(s.keep)
    prm = Ref(params)
    check(ccall((:pnb_hoststep_wcsph_submit, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ref{WcsphParams},
                 Ptr{Cfloat}),
                s.handle, y, v, isnothing(mass) ? C_NULL : mass,
                isnothing(pressure) ? C_NULL : pressure, prm, dv))
    return s
end
# StateEquationCole of the system (benchmarks/smoothed_particle_hydrodynamics.jl:64-69)
function set_state_equation!(s::HostStepper; sound_speed, reference_density, exponent = 1,
                             background_pressure = 0)
    check(ccall((:pnb_hoststep_set_state_equation, libpnb200), Cint,
                (Ptr{Cvoid}, Cfloat, Cfloat, Cfloat, Cfloat),
                s.handle, sound_speed, reference_density, exponent, background_pressure))
    return s
end
Base.wait(s::HostStepper) = (check(ccall((:pnb_hoststep_wait, libpnb200), Cint, (Ptr{Cvoid},), s.handle)); s)

# Any other f: the pairs are found on the device (neighbour list), f runs on the host over the
# exported CSR.  f must only touch host data (scalar indexing of a B200Array is not defined).
function foreach_point_neighbor(f::T, x::B200Array{Float32, 2}, y::B200Array{Float32, 2},
                                nhs::B200GridNeighborhoodSearch;
                                parallelization_backend = default_backend(x),
                                points = axes(x, 2)) where {T}
    lists = NeighborLists(nhs, x, y; sort = false)
    offsets, ids = export_csr(lists)
    pos_diff, distance = pairs(lists, nhs, x, y)
    for i in points, k in (offsets[i] + 1):offsets[i + 1]
        f(i, Int(ids[k]), pos_diff[:, k], distance[k])
    end
    return nothing
end

# ---------------------------------------------------------------------------------------------
# PrecomputedNeighborhoodSearch: device CSR + the reference's DynamicVectorOfVectors layouts
# ---------------------------------------------------------------------------------------------
mutable struct NeighborLists
    handle :: Ptr{Cvoid}
    ndims  :: Int
    function NeighborLists(nhs::B200GridNeighborhoodSearch, x, y; sort = true)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        if eltype(x) === Float64                 # Float64 search: same handle type, same exports
            check(ccall((:pnb_nlist_build_f64, libpnb200), Cint,
                        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint, Ref{Ptr{Cvoid}},
                         Ptr{Cvoid}),
                        nhs.handle, x.ptr, size(x, 2), y.ptr, size(y, 2), sort, ref, C_NULL))
        else
            check(ccall((:pnb_nlist_build_f32, libpnb200), Cint,
                        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint, Ref{Ptr{Cvoid}},
                         Ptr{Cvoid}),
                        nhs.handle, x.ptr, size(x, 2), y.ptr, size(y, 2), sort, ref, C_NULL))
        end
        l = new(ref[], ndims(nhs))
        finalizer(z -> ccall((:pnb_nlist_destroy, libpnb200), Cvoid, (Ptr{Cvoid},), z.handle), l)
        return l
    end
end

n_points(l::NeighborLists) = ccall((:pnb_nlist_n_points, libpnb200), Int64, (Ptr{Cvoid},), l.handle)
n_pairs(l::NeighborLists) = ccall((:pnb_nlist_n_pairs, libpnb200), Int64, (Ptr{Cvoid},), l.handle)

function export_csr(l::NeighborLists)                       # 1-based ids, host arrays
    off = B200Array{Int64}(undef, n_points(l) + 1)
    ids = B200Array{Int32}(undef, max(n_pairs(l), 1))
    check(ccall((:pnb_nlist_export_csr, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                l.handle, off.ptr, ids.ptr, 1, C_NULL))
    return Array(off), Array(ids)[1:n_pairs(l)]
end

# neighbor_lists.backend / .lengths exactly as the reference lays them out
# (src/vector_of_vectors.jl:3-31; transposed parent: :18-23)
function export_dvov(l::NeighborLists, max_neighbors::Integer, transpose_backend::Bool)
    n = n_points(l)
    backend = transpose_backend ? B200Array{Int32}(undef, n, max_neighbors) :
              B200Array{Int32}(undef, max_neighbors, n)
    lengths = B200Array{Int32}(undef, n)
    check(ccall((:pnb_nlist_export_dvov, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Cint, Cint, Ptr{Cvoid}),
                l.handle, backend.ptr, lengths.ptr, max_neighbors, transpose_backend, 1, C_NULL))
    return backend, lengths
end

function pairs(l::NeighborLists, nhs::B200GridNeighborhoodSearch, x, y)
    pd = B200Array{Float32}(undef, l.ndims, max(n_pairs(l), 1))
    d = B200Array{Float32}(undef, max(n_pairs(l), 1))
    check(ccall((:pnb_nlist_pairs_f32, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                l.handle, nhs.handle, x.ptr, y.ptr, pd.ptr, d.ptr, C_NULL))
    return Array(pd), Array(d)
end

mutable struct B200PrecomputedNeighborhoodSearch{NDIMS, ELTYPE, PB} <: AbstractNeighborhoodSearch
    lists               :: Union{Nothing, NeighborLists}
    search_radius       :: ELTYPE
    periodic_box        :: PB
    neighborhood_search :: Union{Nothing, B200GridNeighborhoodSearch{NDIMS}}
    grid_for_sweep      :: B200GridNeighborhoodSearch{NDIMS}   # scalars of the list sweep
    sort_neighbor_lists :: Bool
    max_neighbors       :: Int
    transpose_backend   :: Bool
end

Base.ndims(::B200PrecomputedNeighborhoodSearch{NDIMS}) where {NDIMS} = NDIMS
requires_update(::B200PrecomputedNeighborhoodSearch) = (true, true)    # src/nhs_precomputed.jl:128

function Adapt.adapt_structure(to::B200Backend, nhs::PrecomputedNeighborhoodSearch{NDIMS}) where {NDIMS}
    inner = Adapt.adapt_structure(to, nhs.neighborhood_search)
    return B200PrecomputedNeighborhoodSearch{NDIMS, Float32, typeof(nhs.periodic_box)}(
        nothing, nhs.search_radius, nhs.periodic_box, inner, inner, nhs.sort_neighbor_lists,
        PointNeighbors.max_inner_length(nhs.neighbor_lists, PointNeighbors.max_neighbors(NDIMS)),
        PointNeighbors.transposed_backend(nhs.neighbor_lists))
end

# pushat! errors when a list overflows `max_neighbors` (src/vector_of_vectors.jl:114-121)
function check_list_capacity(s::B200PrecomputedNeighborhoodSearch)
    longest = ccall((:pnb_nlist_max_length, libpnb200), Int64, (Ptr{Cvoid},), s.lists.handle)
    longest > s.max_neighbors && error("cell list is full. Use a larger `max_points_per_cell`.")
    return s
end

# initialize! / update!      src/nhs_precomputed.jl:130-169
function initialize!(s::B200PrecomputedNeighborhoodSearch, x::B200Array{Float32, 2},
                     y::B200Array{Float32, 2}; parallelization_backend = default_backend(x),
                     eachindex_y = axes(y, 2))
    is_all(eachindex_y, size(y, 2)) ||
        error("this neighborhood search does not support inactive points")
    initialize!(s.neighborhood_search, x, y)
    s.lists = NeighborLists(s.neighborhood_search, x, y; sort = s.sort_neighbor_lists)
    return check_list_capacity(s)
end

function update!(s::B200PrecomputedNeighborhoodSearch, x::B200Array{Float32, 2},
                 y::B200Array{Float32, 2}; points_moving = (true, true),
                 parallelization_backend = default_backend(x), eachindex_y = axes(y, 2))
    is_all(eachindex_y, size(y, 2)) ||
        error("this neighborhood search does not support inactive points")
    update!(s.neighborhood_search, x, y; points_moving)
    if any(points_moving)
        s.lists = NeighborLists(s.neighborhood_search, x, y; sort = s.sort_neighbor_lists)
        check_list_capacity(s)
    end
    return s
end

function freeze_neighborhood_search(s::B200PrecomputedNeighborhoodSearch)   # :268-277
    s.neighborhood_search = nothing
    return s
end

struct TLSPHDeformationGradient{A}          # benchmarks/smoothed_particle_hydrodynamics.jl:136-189
    F :: A; current_coordinates :: A
    mass :: Any; material_density :: Any; correction_matrix :: Any
    smoothing_length :: Float32
    kernel_norm :: Float32
end

function foreach_point_neighbor(f::TLSPHDeformationGradient, x::B200Array{Float32, 2},
                                y::B200Array{Float32, 2}, s::B200PrecomputedNeighborhoodSearch;
                                parallelization_backend = default_backend(x),
                                points = axes(x, 2))
    check(ccall((:pnb_tlsph_deformation_grad_f32, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                 Cfloat, Cfloat, Ptr{Cvoid}, Ptr{Cvoid}),
                s.lists.handle, s.grid_for_sweep.handle, x.ptr, f.current_coordinates.ptr,
                f.mass.ptr, f.material_density.ptr, f.correction_matrix.ptr, f.smoothing_length,
                f.kernel_norm, f.F.ptr, C_NULL))
    return nothing
end

# The other TLSPH kernel and the pointwise steps either side of the sweeps
# (benchmarks/smoothed_particle_hydrodynamics.jl:99, 121, 186)
struct TlsphParams                          # pnb_tlsph_params
    smoothing_length :: Float32
    kernel_norm      :: Float32
    young_modulus    :: Float32
    penalty_alpha    :: Float32
end

struct TLSPHInteract{A}                     # TrixiParticles.interact_structure_structure!
    dv :: A; current_coordinates :: A
    mass :: Any; material_density :: Any; pk1_corrected :: Any; deformation_grad :: Any
    params :: TlsphParams
end

function foreach_point_neighbor(f::TLSPHInteract, x::B200Array{Float32, 2},
                                y::B200Array{Float32, 2}, s::B200PrecomputedNeighborhoodSearch;
                                parallelization_backend = default_backend(x),
                                points = axes(x, 2))
    check(ccall((:pnb_tlsph_interact_f32, libpnb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                 Ptr{Cvoid}, Ref{TlsphParams}, Ptr{Cvoid}, Ptr{Cvoid}),
                s.lists.handle, s.grid_for_sweep.handle, x.ptr, f.current_coordinates.ptr,
                f.mass.ptr, f.material_density.ptr, f.pk1_corrected.ptr, f.deformation_grad.ptr,
                Ref(f.params), f.dv.ptr, C_NULL))
    return nothing
end

# TrixiParticles.compute_pk1_corrected!: deformation_grad, correction_matrix, pk1_corrected are
# NDIMS x NDIMS x N device arrays
function compute_pk1_corrected!(pk1_corrected::B200Array{Float32, 3}, deformation_grad::B200Array{Float32, 3},
                                correction_matrix::B200Array{Float32, 3}; young_modulus, poisson_ratio)
    check(ccall((:pnb_tlsph_pk1_corrected_f32, libpnb200), Cint,
                (Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Cfloat, Cfloat, Ptr{Cvoid}, Ptr{Cvoid}),
                size(deformation_grad, 1), size(deformation_grad, 3), deformation_grad.ptr,
                correction_matrix.ptr, young_modulus, poisson_ratio, pk1_corrected.ptr, C_NULL))
    return pk1_corrected
end

# TrixiParticles.compute_pressure! (ContinuityDensity + StateEquationCole): v is (NDIMS + 1) x N
function compute_pressure!(pressure::B200Array{Float32, 1}, v::B200Array{Float32, 2}; sound_speed,
                           reference_density, exponent = 1, background_pressure = 0)
    check(ccall((:pnb_wcsph_compute_pressure_f32, libpnb200), Cint,
                (Cint, Int64, Ptr{Cvoid}, Cfloat, Cfloat, Cfloat, Cfloat, Ptr{Cvoid}, Ptr{Cvoid}),
                size(v, 1) - 1, size(v, 2), v.ptr, sound_speed, reference_density, exponent,
                background_pressure, pressure.ptr, C_NULL))
    return pressure
end

end # module
