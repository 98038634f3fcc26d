# baseline/run_reference.jl -- times the UNMODIFIED RecursiveFactorization.lu! (the real reference) on the host cores.
#
#   julia --project=baseline/_ref -t auto baseline/run_reference.jl N STEPS WARMUP [Float64|Float32]
#
# Not runnable in the build image (no Julia runtime, and the reference's kernels come from un-vendored packages:
# LoopVectorization, TriangularSolve, Polyester, StrideArraysCore -- SURVEY.md F2/F3).  It is staged for the day a
# driver provides Julia plus an instantiated environment under baseline/_ref (git-ignored): bench.py --impl reference
# calls it when `julia` is on PATH and this environment loads, and labels the line `"kind": "reference"`;
# otherwise it times the C port (oracle/rf_oracle.c) and says `"kind": "port"`.
#
# Same workload definition as bench.py: A[i,j] ~ U[0,1), column-major, partial pivoting, threaded (`Val(true)`,
# src/lu.jl:132-144), `check = false`; prints ONE JSON line on stdout.
import RecursiveFactorization
using LinearAlgebra, Random

function main(args)
    n = parse(Int, get(args, 1, "16384"))
    steps = parse(Int, get(args, 2, "5"))
    warmup = parse(Int, get(args, 3, "3"))
    T = get(args, 4, "Float64") == "Float32" ? Float32 : Float64
    Random.seed!(12)                                   # test/runtests.jl:9
    A0 = rand(T, n, n)                                 # test/runtests.jl:45
    A = similar(A0)
    ipiv = Vector{LinearAlgebra.BlasInt}(undef, n)
    times = Float64[]
    info = 0
    for it in 1:(warmup + steps)
        copyto!(A, A0)
        t = @elapsed begin
            F = RecursiveFactorization.lu!(A, ipiv, Val(true), Val(true); check = false)   # src/lu.jl:97-130
            info = F.info
        end
        it > warmup && push!(times, t)
    end
    ms = 1e3 * sum(times) / length(times)
    gflops = (2.0 * n^3 / 3.0) / (ms * 1e-3) / 1e9
    # residual of the last factorization: ||PA - LU||_F / ||A||_F through 4 random +-1 probes (O(n^2))
    p = collect(1:n)
    for i in 1:n
        j = ipiv[i]
        p[i], p[j] = p[j], p[i]
    end
    x = rand([-1.0, 1.0], n, 4)
    L = UnitLowerTriangular(A); U = UpperTriangular(A)
    r = norm(A0[p, :] * x - L * (U * x)) / sqrt(4) / norm(A0)
    println("{\"impl\": \"reference\", \"kind\": \"reference\", \"n\": $n, \"eltype\": \"$T\", \"steps\": $steps, " *
            "\"warmup\": $warmup, \"ms_per_step\": $ms, \"value\": $gflops, \"unit\": \"GFLOP/s\", " *
            "\"threads\": $(Threads.nthreads()), \"info\": $info, \"residual_fro_rel_est\": $r, " *
            "\"version\": \"$(pkgversion(RecursiveFactorization))\"}")
end

main(ARGS)
