# CorleoneB200.jl -- reference-side binding of libcorleone_b200.so (include/corleone_b200.h).
#
# This file is what a Corleone.jl maintainer adds (e.g. as ext/CorleoneB200Extension.jl) to route the
# shooting hot path through the B200 library.  It uses only `ccall`: no CUDA.jl, no kernel codegen.
# Julia is not available in the build container, so this file is written against the reference sources
# and the C header and is NOT executed by the test-suite; the same ABI is exercised from Python (ctypes,
# corleone.jl_b200/native.py) by tests/test_cabi.py and tests/test_gpu_parity.py.
#
# Hooks used (all exist in the reference):
#   * `ensemble_alg` field of MultipleShootingLayer (src/multiple_shooting.jl:19,71) -- a new
#     SciMLBase.EnsembleAlgorithm subtype selects the device path;
#   * `Corleone._parallel_solve(shooting, u0, ps, st)` (src/multiple_shooting.jl:118-130) is overloaded for
#     that subtype and hands ALL intervals to one cb200_eval;
#   * `OptimizationFunction(...; kwargs...)` forwards `grad=` / `cons_j=` (src/dynprob.jl:143-164).
module CorleoneB200

using Corleone
using Corleone: MultipleShootingLayer, SingleShootingLayer, Trajectory
using LuxCore, SciMLBase

const LIB = get(ENV, "CORLEONE_B200_LIB", "libcorleone_b200.so")

# ---------------------------------------------------------------- C structs (include/corleone_b200.h)
struct Desc                      # cb200_desc
    struct_size::Int32
    model::Int32
    method::Int32
    steps_per_piece::Int32
    nx::Int32
    np_all::Int32
    n_controls::Int32
    n_tunable_p::Int32
    n_intervals::Int32
    n_quadrature::Int32
    device::Int32
    fim_offset::Int32
    fim_n::Int32
    fim_mode::Int32                  # cb200_fim_mode
    u0::Ptr{Float64}
    tunable_p::Ptr{Int64}
    control_indices::Ptr{Int64}
    quadrature_indices::Ptr{Int64}
    n_pieces::Ptr{Int32}
    tspans::Ptr{Float64}
    index_grid::Ptr{Int64}
    n_u0::Ptr{Int32}
    n_local_controls::Ptr{Int32}
    is_shooting::Ptr{Int32}
    ic_sorting::Ptr{Int64}
    ic_n_constants::Ptr{Int32}
    ic_constants::Ptr{Int64}
    n_shooting_indices::Ptr{Int32}
    shooting_indices::Ptr{Int64}
    model_data::Ptr{Float64}
    model_data_len::Int64
    fim_base_nx::Int32
    n_observed::Int32
    n_observation_rows::Ptr{Int32}
    observation_grid::Ptr{Int64}
    abstol::Float64
    reltol::Float64
    saveat::Ptr{Float64}             # problem.kwargs saveat (sorted), or C_NULL  -- version 201
    n_saveat::Int32
    reserved_desc::Int32
end

struct Out                       # cb200_out
    xend::Ptr{Float64}
    sens::Ptr{Float64}
    traj::Ptr{Float64}
    defects_kind::Ptr{Float64}
    defects_stage::Ptr{Float64}
    objective::Ptr{Float64}
    grad::Ptr{Float64}
    fim::Ptr{Float64}
    fim_sum::Ptr{Float64}
    objective_sum::Ptr{Float64}
    grad_sum::Ptr{Float64}
    fim_init::Ptr{Float64}
    objective_row::Int32
    reserved::Int32
    jac_vals::Ptr{Float64}
    criteria::Ptr{Float64}
    grad_compact::Ptr{Float64}
    traj_sens::Ptr{Float64}
    status::Ptr{Int32}
    defects_gathered::Ptr{Float64}   # version 200: CB200_COMM_GATHER
    comm_B_total::Int64
    comm_b0::Int64
    resident::UInt32                 # want-bits of results that stay on the device (cb200_resident)
    reserved2::UInt32
end

struct CommInfo                  # cb200_comm_info
    nranks::Int32
    rank::Int32
    transport::Int32
    reserved::Int32
    epoch::Int64
    max_B_total::Int64
    error::Int64
end

struct Sizes                     # cb200_sizes
    nvars::Int64
    n_defects::Int64
    n_samples::Int64
    n_row::Int64
    n_sens::Int64
    n_units::Int64
    jac_nnz::Int64
    jac_nnz_var::Int64
    n_traj_sens::Int64
end

const WANT_XEND, WANT_SENS, WANT_TRAJ, WANT_DEFECTS = 0x001, 0x002, 0x004, 0x008
const WANT_OBJ, WANT_GRAD, WANT_FIM, WANT_JAC, WANT_CRIT = 0x010, 0x020, 0x040, 0x080, 0x100
const WANT_GRAD_COMPACT, WANT_TRAJ_SENS = 0x200, 0x400
const MEM_DEVICE, PS_SOA, OUT_SOA, COMM_REDUCE, COMM_GATHER = 0x1, 0x2, 0x4, 0x8, 0x10
const COMM_ID_BYTES = 128

"""
Layout self-check: the structs above restate include/corleone_b200.h by hand; the library reports sizeof and the offsets
of the trailing fields of each (cb200_abi_layout) and a drift stops here instead of corrupting memory.
"""
function check_abi()
    layout(which) = (v = zeros(Int64, 16); n = ccall((:cb200_abi_layout, LIB), Cint, (Int32, Ptr{Int64}, Int32), which, v, 16); v[1:n])
    want = (
        (0, Desc, (:u0, :model_data_len, :observation_grid, :reltol, :saveat, :n_saveat)),
        (1, Out, (:objective_row, :jac_vals, :status, :defects_gathered, :comm_b0, :resident)),
        (2, Sizes, (:n_traj_sens,)),
        (3, CommInfo, (:epoch, :error)),
    )
    for (which, T, fields) in want
        mine = vcat(sizeof(T), [fieldoffset(T, findfirst(==(f), fieldnames(T))) for f in fields])
        theirs = layout(which)
        mine == theirs || error("CorleoneB200: struct $(T) does not match libcorleone_b200 (header version " *
                                "$(ccall((:cb200_version, LIB), Cint, ()))): $mine vs $theirs")
    end
    return true
end

const MODELS = Dict(:lotka3 => 0, :lqr => 1, :lotka2 => 2, :lotka2_oed => 3, :lin1d => 4, :lin1d_oed => 5,
                    :dense20 => 6, :egerstedt => 7, :vdp3 => 8, :cartpole5 => 9, :rocket3 => 10, :quadrotor7 => 11, :threetank4 => 12,
                    :batchreactor2 => 13, :lotkacomp3 => 14, :f8aircraft4 => 15, :doubleosc6 => 16,
                    :lotka2_oed_cu => 17, :lotka2_oed_cf => 18, :lotka2_oed_ds => 19, :lotka2_oed_du => 20,
                    :lin1d_oed_cf => 21, :lin1d_oed_ds => 22,
                    :brysondenham3 => 23, :catalystmixing2 => 24, :cushioned3 => 25, :denbigh3 => 26, :dielectrophoretic3 => 27, :electriccar4 => 28, :fuller3 => 29, :jackson3 => 30, :mountaincar3 => 31, :particlesteering5 => 32, :raomease2 => 33, :robbins4 => 34, :moonlanding3 => 35, :lotkashared4 => 36, :hangingchain3 => 37, :tubularreactor2 => 38, :ocean4 => 39, :ductedfan7 => 40, :hangglider4 => 41,
                    :dense16 => 42, :dense24 => 43, :dense32 => 44)

lasterror() = unsafe_string(ccall((:cb200_last_error, LIB), Cstring, ()))
check(rc, what) = rc == 0 || error("$what failed ($rc): $(lasterror())")

# ---------------------------------------------------------------- the ensemble algorithm that selects the GPU
"""
    EnsembleB200(; model, steps_per_piece = 1, method = :tsit5, device = 0, model_data = Float64[])

`MultipleShootingLayer(prob, Tsit5(), tpoints...; ensemble_alg = EnsembleB200(model = :lotka3, steps_per_piece = 4))`
integrates every shooting interval on the GPU with the fixed step `(t1 - t0)/steps_per_piece` per control
piece -- the `adaptive = false, dt = h` mode of `solve`.  `model` names the right-hand side compiled into
the library (the problem's Julia closure cannot cross a C ABI).
"""
mutable struct EnsembleB200 <: SciMLBase.EnsembleAlgorithm
    model::Symbol
    steps_per_piece::Int
    method::Symbol
    device::Int
    model_data::Vector{Float64}
    fim::Tuple{Int, Int}                 # (1-based offset of the first F_triu state, n) or (0, 0)
    fim_mode::Int                        # 0 end state, 1 Discrete, 2 DiscreteSampled, 3 fixed ContinuousSampled
    fim_base_nx::Int                     # states of the un-augmented problem (sample-based read-outs)
    n_observed::Int
    handle::Ptr{Cvoid}
    sizes::Sizes
    nx::Int
    abi_checked::Bool
end
function EnsembleB200(; model::Symbol, steps_per_piece = 1, method = :tsit5, device = 0,
        model_data = Float64[], fim = (0, 0), fim_mode = 0, fim_base_nx = 0, n_observed = 0)
    e = EnsembleB200(model, steps_per_piece, method, device, collect(Float64, model_data), fim, fim_mode,
        fim_base_nx, n_observed, C_NULL,
        Sizes(0, 0, 0, 0, 0, 0, 0, 0, 0), 0, false)
    finalizer(e) do x
        x.handle == C_NULL || ccall((:cb200_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle)
    end
    return e
end

"Restrict the following evaluations to the shooting intervals `r` (interval sharding across devices, B < #GPUs)."
set_interval_range!(e::EnsembleB200, r::UnitRange{Int}) =
    check(ccall((:cb200_set_interval_range, LIB), Cint, (Ptr{Cvoid}, Int32, Int32), e.handle, first(r), last(r)),
        "cb200_set_interval_range")

flat_tspans(ts::Tuple{Vararg{Tuple{<:Real, <:Real}}}) = collect(ts)
flat_tspans(ts::Tuple) = reduce(vcat, map(flat_tspans, ts))               # <=100-piece chunks (local_controls.jl:190-193)
flat_grid(g::AbstractMatrix) = g
flat_grid(g::AbstractVector) = reshape(g, :, 1)
flat_grid(g::Tuple) = reduce(hcat, map(flat_grid, g))                     # (local_controls.jl:173-176)

"Build the handle once from the layer's static state `st` (src/single_shooting.jl:343-351), as Julia holds it."
function handle!(e::EnsembleB200, shooting::MultipleShootingLayer, ps, st::NamedTuple{fields}) where {fields}
    e.handle != C_NULL && return e.handle
    layer = shooting.layer
    prob = layer.problem
    nx = length(prob.u0)
    np_all = length(prob.p)
    cidx = collect(Int64, layer.control_indices)
    tunable_p = collect(Int64, setdiff(1:np_all, cidx))
    quad = collect(Int64, Corleone.get_quadrature_indices(layer))
    I = length(fields)
    n_pieces = Int32[]; tsp = Float64[]; grid = Int64[]; n_u0 = Int32[]; n_c = Int32[]; is_sh = Int32[]
    sorting = Int64[]; ncon = Int32[]; cons = Int64[]; nsi = Int32[]; sidx = Int64[]
    for (i, f) in enumerate(fields)
        sti, psi = getproperty(st, f), getproperty(ps, f)
        ts = flat_tspans(sti.tspans)
        push!(n_pieces, length(ts))
        for (a, b) in ts
            push!(tsp, a, b)
        end
        append!(grid, vec(flat_grid(sti.index_grid)))                      # column-major, 1-based
        push!(n_u0, length(psi.u0)); push!(n_c, length(psi.controls)); push!(is_sh, i > 1)
        append!(sorting, sti.initial_condition.sorting)
        push!(ncon, length(sti.initial_condition.constants)); append!(cons, sti.initial_condition.constants)
        push!(nsi, length(sti.shooting_indices)); append!(sidx, sti.shooting_indices)
    end
    u0 = collect(Float64, prob.u0)
    md = e.model_data
    # WeightedObservation.grid of every interval (lib/CorleoneOED/src/oed.jl:2-16,307-323,350-366), rows flattened
    n_obs_rows = Int32[]; obs_grid = Int64[]
    for f in fields
        sti = getproperty(st, f)
        haskey(sti, :observation_grid) || continue
        push!(n_obs_rows, length(sti.observation_grid.grid))
        foreach(row -> append!(obs_grid, row), sti.observation_grid.grid)
    end
    # problem.kwargs saveat (the discrete OED layers put their measurement times there, lib/CorleoneOED/src/oed.jl:106-113)
    saveat = haskey(prob.kwargs, :saveat) ? sort!(collect(Float64, prob.kwargs[:saveat])) : Float64[]
    e.abi_checked || (check_abi(); e.abi_checked = true)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve u0 tunable_p cidx quad n_pieces tsp grid n_u0 n_c is_sh sorting ncon cons nsi sidx md n_obs_rows obs_grid saveat begin
        d = Desc(sizeof(Desc), MODELS[e.model], e.method === :tsit5 ? 0 : (e.method === :rk4 ? 1 : 2), e.steps_per_piece, nx, np_all,
            length(cidx), length(tunable_p), I, length(quad), e.device, e.fim[1], e.fim[2], e.fim_mode,
            pointer(u0), pointer(tunable_p), pointer(cidx), pointer(quad), pointer(n_pieces), pointer(tsp),
            pointer(grid), pointer(n_u0), pointer(n_c), pointer(is_sh), pointer(sorting), pointer(ncon),
            pointer(cons), pointer(nsi), pointer(sidx), isempty(md) ? Ptr{Float64}(C_NULL) : pointer(md),
            length(md), e.fim_base_nx, e.n_observed,
            isempty(n_obs_rows) ? Ptr{Int32}(C_NULL) : pointer(n_obs_rows),
            isempty(obs_grid) ? Ptr{Int64}(C_NULL) : pointer(obs_grid),
            Float64(get(prob.kwargs, :abstol, 1.0e-6)), Float64(get(prob.kwargs, :reltol, 1.0e-3)),
            isempty(saveat) ? Ptr{Float64}(C_NULL) : pointer(saveat), length(saveat), 0)
        check(ccall((:cb200_create, LIB), Cint, (Ref{Desc}, Ref{Ptr{Cvoid}}), d, h), "cb200_create")
    end
    e.handle = h[]
    s = Ref(Sizes(0, 0, 0, 0, 0, 0, 0, 0, 0))
    check(ccall((:cb200_get_sizes, LIB), Cint, (Ptr{Cvoid}, Ref{Sizes}), e.handle, s), "cb200_get_sizes")
    e.sizes = s[]
    e.nx = nx
    return e.handle
end

"flat candidate in ComponentArray order: intervals in order, fields (u0, p, controls) (ext/CorleoneComponentArraysExtension.jl:15-23)"
flatten(ps::NamedTuple) = reduce(vcat, (vcat(p.u0, p.p, p.controls) for p in values(ps)))

"""
    evaluate(e, P; want, objective_row = 1)   (after `handle!(e, shooting, ps, st)`)

Batched evaluation -- the axis the reference does not have.  `P` is an `nvars x B` Matrix, one flat
candidate per column.  Returns a NamedTuple of `n x B` matrices (host memory, owned by Julia).
"""
function evaluate(e::EnsembleB200, P::Matrix{Float64}; want::Integer, objective_row::Integer = 1,
        fim_init::Union{Nothing, Matrix{Float64}} = nothing)
    s = e.sizes
    B = size(P, 2)
    alloc(flag, n) = (want & flag) != 0 ? Matrix{Float64}(undef, n, B) : Matrix{Float64}(undef, 0, 0)
    xend = alloc(WANT_XEND, e.nx * s.n_units)
    sens = alloc(WANT_SENS, s.n_sens)
    traj = alloc(WANT_TRAJ, s.n_row * s.n_samples)
    dk = alloc(WANT_DEFECTS, s.n_defects)
    ds = alloc(WANT_DEFECTS, s.n_defects)
    obj = alloc(WANT_OBJ, 1)
    grad = alloc(WANT_GRAD, s.nvars)
    nf = e.fim[2]
    fim = alloc(WANT_FIM, nf * nf)
    jac = alloc(WANT_JAC, s.jac_nnz_var)
    crit = alloc(WANT_CRIT, 6)
    status = zeros(Int32, B)
    fsum = zeros(nf, nf)
    ptr(a) = isempty(a) ? Ptr{Float64}(C_NULL) : pointer(a)
    GC.@preserve P xend sens traj dk ds obj grad fim fsum fim_init jac crit status begin
        o = Out(ptr(xend), ptr(sens), ptr(traj), ptr(dk), ptr(ds), ptr(obj), ptr(grad), ptr(fim),
            (want & WANT_FIM) != 0 ? pointer(fsum) : Ptr{Float64}(C_NULL), C_NULL, C_NULL,
            fim_init === nothing ? Ptr{Float64}(C_NULL) : pointer(fim_init), objective_row, 0, ptr(jac), ptr(crit), Ptr{Float64}(C_NULL), Ptr{Float64}(C_NULL), pointer(status),
            Ptr{Float64}(C_NULL), 0, 0, UInt32(0), UInt32(0))
        check(ccall((:cb200_eval, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, UInt32, UInt32, Ref{Out}),
                e.handle, P, B, size(P, 1), want, 0, o), "cb200_eval")
    end
    return (; xend, sens, traj, defects_kind = dk, defects_stage = ds, objective = obj, grad, fim, fim_sum = fsum, jac_vals = jac, criteria = crit, status)
end

# ---------------------------------------------------------------- drop-in: layer(nothing, ps, st) on the GPU
# Replaces mythreadmap over intervals (src/Corleone.jl:22-24) AND the Trajectory merge
# (src/multiple_shooting.jl:139-182): the library returns the merged samples, the stage-ordered defects and
# the quadrature running sums in one call.
function (shooting::MultipleShootingLayer{L, I, EnsembleB200})(u0, ps, st::NamedTuple{fields}) where {L, I, fields}
    e = shooting.ensemble_alg
    handle!(e, shooting, ps, st)
    r = evaluate(e, reshape(flatten(ps), :, 1); want = WANT_TRAJ | WANT_DEFECTS)
    s = e.sizes
    U = reshape(r.traj, Int(s.n_row), Int(s.n_samples))
    u = [U[:, k] for k in axes(U, 2)]
    # time points: piece boundaries and saveat times, duplicate interval boundaries dropped (:145-147)
    t = Vector{Float64}(undef, Int(s.n_samples))
    check(ccall((:cb200_sample_times, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.handle, t), "cb200_sample_times")
    # split the stage-ordered defect vector into the (u0, p, controls) NamedTuple per stage (:149-169)
    nx = Corleone.statelength(getproperty(st, fields[1]).initial_condition)
    d, o = vec(r.defects_stage), 0
    vals = map(enumerate(fields)) do (i, f)
        i == 1 && return (u0 = Float64[], p = Float64[], controls = Float64[])
        idx = getproperty(st, f).shooting_indices
        nu, nc, npar = count(<=(nx), idx), count(>(nx), idx), length(getproperty(ps, f).p)
        v = (u0 = d[(o + 1):(o + nu)], p = d[(o + nu + 1):(o + nu + npar)],
            controls = d[(o + nu + npar + 1):(o + nu + npar + nc)])
        o += nu + npar + nc
        v
    end
    offsets = cumsum([length(flat_tspans(getproperty(st, f).tspans)) + 1 for f in fields[1:(end - 1)]])
    sys = getproperty(st, fields[1]).symcache
    return Trajectory(sys, u, getproperty(ps, fields[1]).p, t, NamedTuple{fields}(Tuple(vals)), offsets), st
end

# ---------------------------------------------------------------- NLP callbacks without AD re-integration
# OptimizationProblem(layer, AutoForwardDiff(), Val(:ComponentArrays); loss = :x₃,
#                     grad = b200_grad(e, row), cons_j = b200_cons_j(e), cons_jac_prototype = b200_jac_prototype(e))
# (kwargs are forwarded to OptimizationFunction, src/dynprob.jl:161).
b200_grad(e::EnsembleB200, objective_row) = (g, x, st) ->
    (g .= vec(evaluate(e, reshape(collect(x), :, 1); want = WANT_GRAD, objective_row).grad); g)

function jac_structure(e::EnsembleB200)
    n = Int(e.sizes.jac_nnz)
    rows, cols, src = Vector{Int64}(undef, n), Vector{Int64}(undef, n), Vector{Int64}(undef, n)
    check(ccall((:cb200_jac_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
            e.handle, rows, cols, src), "cb200_jac_structure")
    return rows, cols, src
end

function b200_cons_j(e::EnsembleB200)
    rows, cols, src = jac_structure(e)
    nv = Int(e.sizes.jac_nnz_var)                        # the leading triplets carry sensitivity values
    return function (J, x, st)                          # rows of the shooting block follow the point constraints
        vals = vec(evaluate(e, reshape(collect(x), :, 1); want = WANT_JAC).jac_vals)
        fill!(J, 0)
        for k in 1:nv
            J[rows[k], cols[k]] += vals[k]               # assembled on the device (K5), triplet order
        end
        for k in (nv + 1):length(rows)
            J[rows[k], cols[k]] += src[k] == -1 ? 1.0 : -1.0   # matching conditions w.r.t. the next interval's variables
        end
        return J
    end
end

# ---------------------------------------------------------------- multi-GPU data plane (cb200_comm_*)
# One Julia process (or task) per GPU; replaces mythreadmap(::EnsembleDistributed) = pmap over intervals (src/Corleone.jl:24).
# Rank 0 creates the id and hands it to the others (Distributed.jl `remotecall_fetch`, MPI.bcast, ...); after comm_init!
# an evaluation with COMM_REDUCE / COMM_GATHER finishes with the all-reduced sums and the all-gathered defects, written by
# the evaluation's own kernels over NVLink -- no collective call on the host.
function comm_unique_id()
    id = zeros(UInt8, COMM_ID_BYTES)
    check(ccall((:cb200_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id), "cb200_comm_unique_id")
    return id
end
comm_init!(e::EnsembleB200, nranks::Integer, rank::Integer, id::Vector{UInt8}; max_B_total::Integer = 0) =
    check(ccall((:cb200_comm_init, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}, Int64), e.handle, nranks, rank, id,
            max_B_total), "cb200_comm_init")
comm_destroy!(e::EnsembleB200) = check(ccall((:cb200_comm_destroy, LIB), Cint, (Ptr{Cvoid},), e.handle), "cb200_comm_destroy")
function comm_info(e::EnsembleB200)
    r = Ref(CommInfo(0, 0, 0, 0, 0, 0, 0))
    check(ccall((:cb200_comm_query, LIB), Cint, (Ptr{Cvoid}, Ref{CommInfo}), e.handle, r), "cb200_comm_query")
    return r[]
end

"""
    evaluate_sharded(e, P, b0, B_total; objective_row)

This rank's shard `P` (`nvars x B`, global candidates `b0 + 1 : b0 + B` of `B_total`) of a batch-sharded evaluation:
returns the objective and gradient sums over ALL ranks' candidates (bit-identical on every rank) and the shooting
constraints of all `B_total` candidates (`n_defects x B_total`).  Collective: every rank calls it.
"""
function evaluate_sharded(e::EnsembleB200, P::Matrix{Float64}, b0::Integer, B_total::Integer; objective_row::Integer = 1)
    s = e.sizes
    B = size(P, 2)
    dk = Matrix{Float64}(undef, s.n_defects, B)
    gathered = Matrix{Float64}(undef, s.n_defects, B_total)
    obj = Matrix{Float64}(undef, 1, B)
    osum = zeros(1)
    gsum = zeros(s.nvars)
    want = WANT_DEFECTS | WANT_OBJ | WANT_GRAD
    N = Ptr{Float64}(C_NULL)
    GC.@preserve P dk gathered obj osum gsum begin
        o = Out(N, N, N, pointer(dk), N, pointer(obj), N, N, N, pointer(osum), pointer(gsum), N, objective_row, 0, N, N, N, N,
            Ptr{Int32}(C_NULL), pointer(gathered), B_total, b0, UInt32(0), UInt32(0))
        check(ccall((:cb200_eval, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, UInt32, UInt32, Ref{Out}),
                e.handle, P, B, size(P, 1), want, COMM_REDUCE | COMM_GATHER, o), "cb200_eval")
    end
    return (; objective = obj, objective_sum = osum[1], grad_sum = gsum, defects_kind = dk, defects_gathered = gathered)
end

end # module
