# gen_reference_fixtures.jl -- run the UNMODIFIED reference (DarianNwankwo/Rollout-Bayesian-Optimization) on two small problems and
# dump everything the hot path consumed and produced, so that the CPU oracle (oracle/rbo_oracle.cpp) and the CUDA path can be
# replayed against REAL reference outputs (tests/test_reference_fixtures.py). This is the road from "parity unpinned" to pinned:
# Julia is not installed in the build environment, so this script has never been executed there; the first person with Julia
# and the reference's dependencies (Optim, Sobol, Distributions, ForwardDiff, FastGaussQuadrature, IterTools, Plots, NPZ) runs
#
#     ROLLOUT_BO_REFERENCE_DIR=/path/to/Rollout-Bayesian-Optimization julia oracle/julia/gen_reference_fixtures.jl tests/golden
#
# and commits tests/golden/julia_*.npz. Nothing of the reference is copied: its files are included from the user's checkout.
#
# What is recorded per fixture (all Float64, Julia's column-major layout):
#   inputs : X (d x N), L (N x N lower), y, c = cs[1], sigma_n2, ell (Matern-5/2), x0, theta, lbs, ubs, h, fmini = minimum(s.y) over the
#            zero-padded capacity vector (rollout.jl:109,234), rn = tp.rnstream_sequence (M x (d+1) x (h+1)), starts (d x (S+2)),
#            dual_dirs (d x h x M): the rand(dim) draws of solve_dual_y (rollout.jl:133) in the order the reference consumed them
#   outputs: xs (d x (h+1) x M) -- the IPNewton x-path (fs.X[:, N+1:N+h+1] after rollout!, rollout.jl:39-74) --, ys, gys (the sampled
#            observations / gradients, observables.jl:117-118), values = resolutions, grad_x, grad_theta (rollout.jl:233-277),
#            best_index (0-based t of rollout.jl:235), grad_case (1/2/3 of rollout.jl:239-251)
# The replay is TEACHER-FORCED on xs: everything except Optim's iterates is then pinned to the real reference.
using Random, LinearAlgebra
import NPZ

const REF = get(ENV, "ROLLOUT_BO_REFERENCE_DIR", "")
isdir(REF) || error("set ROLLOUT_BO_REFERENCE_DIR to a checkout of the reference")
include(joinpath(REF, "rollout_bayesian_optimization.jl"))

# ---- record the rand(dim) stream of solve_dual_y (rollout.jl:133) without touching the reference: the global RNG is seeded,
# the trajectory gradient is computed, and the same seed then replays the draws in consumption order
function gradient_with_recorded_directions(T::Trajectory, d::Int, h::Int; seed::Int)
    Random.seed!(seed)
    ∇x, ∇θ = gradient(T)                                          # rollout.jl:233-277
    fmini = minimum(get_observations(get_base_surrogate(T)))
    t, bestpt = best(T)                                           # rollout.jl:85-105: (0-based index, (x = .., y = ..))
    dirs = zeros(d, max(h, 1))
    case = fmini <= bestpt.y ? 1 : (t == 0 ? 2 : 3)               # rollout.jl:239-251
    if case == 3
        Random.seed!(seed)
        for j in t:-1:1                                           # rollout.jl:259-262: solve_dual_y(solve_index = j-1) draws rand(dim)
            dirs[:, j] = rand(d)
        end
    end
    return ∇x, ∇θ, dirs, t, case
end

function fixture(name; testfn, d, N, h, M, S, ell, seed)
    Random.seed!(seed)
    lbs, ubs = get_bounds(testfn)
    X = lbs .+ (ubs .- lbs) .* rand(d, N)
    y = [testfn(X[:, j]) for j in 1:N]
    sur = Surrogate(Matern52([ell]), X, y; capacity = N + h + 8, decision_rule = EI(), σn2 = 1e-6)
    fs = FantasySurrogate(sur, h)
    x0 = (lbs .+ ubs) ./ 2
    θ = [0.0]
    T = Trajectory(sur, fs; start = x0, hypers = θ, horizon = h)
    tp = TrajectoryParameters(start = x0, hypers = θ, horizon = h, mc_iterations = M, use_low_discrepancy_sequence = true,
                              spatial_lowerbounds = lbs, spatial_upperbounds = ubs)
    starts = generate_initial_guesses(S, lbs, ubs)
    xs = zeros(d, h + 1, M); ys = zeros(h + 1, M); gys = zeros(d, h + 1, M)
    values = zeros(M); gx = zeros(d, M); gθ = zeros(length(θ), M); dual = zeros(d, max(h, 1), M)
    best_index = zeros(Int32, M); grad_case = zeros(Int32, M)
    Nobs = get_known_observations(fs)
    for m in 1:M
        set_start!(T, x0)
        obs = StochasticObservable(fantasy_surrogate = fs, stdnormal = get_samples_rnstream(tp, sample_index = m), max_invocations = h + 1)  # rollout.jl:295-300
        attach_observable!(T, obs)
        rollout!(T, lowerbounds = lbs, upperbounds = ubs, get_observation = get_observable(T), xstarts = starts)                 # rollout.jl:309
        xs[:, :, m] = fs.X[:, Nobs+1:Nobs+h+1]
        ys[:, m] = obs.observations[1:h+1]
        gys[:, :, m] = obs.gradients[:, 1:h+1]
        values[m] = resolve(T)                                                                                                    # rollout.jl:318
        ∇x, ∇θ, dirs, t, case = gradient_with_recorded_directions(T, d, h; seed = 1906 + m)
        gx[:, m] = ∇x; gθ[:, m] = ∇θ; dual[:, :, m] = dirs; best_index[m] = t; grad_case[m] = case
        reset!(fs)                                                                                                                 # rollout.jl:325
    end
    NPZ.npzwrite(joinpath(ARGS[1], "julia_$(name).npz"), Dict(
        "X" => Matrix(sur.X[:, 1:N]), "L" => Matrix(sur.L[1:N, 1:N]), "y" => sur.y[1:N], "c" => sur.c[1:N], "sigma_n2" => 1e-6, "ell" => ell,
        "x0" => x0, "theta" => θ, "lbs" => lbs, "ubs" => ubs, "h" => h, "fmini" => minimum(get_observations(sur)),
        "rn" => tp.rnstream_sequence, "starts" => starts, "dual_dirs" => dual, "xs" => xs, "ys" => ys, "gys" => gys, "values" => values,
        "grad_x" => gx, "grad_theta" => gθ, "best_index" => best_index, "grad_case" => grad_case))
    println("wrote julia_$(name).npz: cases ", [count(==(c), grad_case) for c in 1:3])
end

length(ARGS) == 1 || error("usage: julia gen_reference_fixtures.jl <output dir>")
fixture("branin_h1"; testfn = TestBraninHoo(), d = 2, N = 10, h = 1, M = 32, S = 8, ell = 1.0, seed = 1906)     # BASELINE config C1 shape
fixture("hartmann6_h2"; testfn = TestHartmann6D(), d = 6, N = 20, h = 2, M = 16, S = 8, ell = 0.5, seed = 1907)
