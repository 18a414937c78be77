# baseline/run_reference.jl -- times the REAL reference (SciML/Corleone.jl) on the bench box's host cores and dumps
# fixed-step vectors for parity, when a Julia installation with the reference's dependency closure exists there.
# Neither exists in the build image (no `julia`, no network; OrdinaryDiffEq / ForwardDiff are not vendored), so this script
# is committed unexecuted; bench.py's `--impl reference` arm falls back to the C restatement (kind = "port").
#
#   julia --project=/path/to/Corleone.jl --threads auto baseline/run_reference.jl [out.json]
#
# What it does: BASELINE config 2 (LQR, 100 shooting intervals, control step 0.03) with `adaptive = false, dt = h`
# (h = 0.015: K = 2 fixed steps per control piece) through the reference's own layer call and ForwardDiff, for a sample of
# the candidate batch; prints units/s (one unit = one (interval, candidate) solve with sensitivities) and writes
# end states / defects / objective / gradient of the first candidates for tests/golden/.
using Corleone, OrdinaryDiffEqTsit5, ForwardDiff, ComponentArrays, LuxCore, Random, JSON

function lqr!(du, u, p, t)
    a, b, uc = p
    du[1] = a * u[1] + b * uc
    du[2] = 10 * (u[1] - 3)^2 + 0.1 * uc^2
    return nothing
end

function main(out)
    I, nctl, K = 100, 400, 2
    tgrid = [3.0 * k / 100.0 for k in 0:(nctl - 1)]
    tspan = (0.0, 3.0 * nctl / 100.0)
    prob = ODEProblem(lqr!, [1.0, 0.0], tspan, [-1.0, 1.0, 0.0]; adaptive = false, dt = 0.03 / K)
    control = ControlParameter(tgrid; name = :u, controls = zeros(nctl))
    tpoints = [3.0 * (4 * k) / 100.0 for k in 0:(I - 1)]
    layer = MultipleShootingLayer(prob, Tsit5(), tpoints...; controls = (3 => control,), quadrature_indices = [2],
        ensemble_alg = EnsembleThreads())
    rng = Random.Xoshiro(20261019 + 2)
    ps, st = LuxCore.setup(rng, layer)
    x0 = collect(ComponentArray(ps))
    ax = getaxes(ComponentArray(ps))
    constraints(x) = begin
        traj, _ = layer(nothing, ComponentArray(x, ax), st)
        res = zeros(eltype(x), Corleone.get_number_of_shooting_constraints(layer))
        Corleone.shooting_constraints!(res, traj)
        vcat(last(traj.u)[2], res)
    end
    B = max(Threads.nthreads(), 16)
    X = [x0 .+ 0.1 .* randn(rng, length(x0)) for _ in 1:B]
    constraints(X[1]); ForwardDiff.jacobian(constraints, X[1])          # compile
    t = @elapsed for x in X
        ForwardDiff.jacobian(constraints, x)
    end
    units = I * B
    println("reference (Corleone.jl, $(Threads.nthreads()) threads): $(units / t) units/s over $B candidates")
    open(out, "w") do io
        JSON.print(io, Dict("units_per_s" => units / t, "threads" => Threads.nthreads(), "candidates" => B,
            "x" => X[1:2], "values" => [constraints(x) for x in X[1:2]]))
    end
end

main(length(ARGS) >= 1 ? ARGS[1] : "reference_fixed_step.json")
