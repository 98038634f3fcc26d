# tests/golden/make_golden_julia.jl -- regenerate golden OUTPUTS from the real RecursiveFactorization.jl.
#
#   python tests/golden/export_inputs.py tests/golden/julia_io          # writes <case>.in (raw column-major) + index.txt
#   julia --project=baseline/_ref tests/golden/make_golden_julia.jl tests/golden/julia_io
#   python tests/golden/export_inputs.py --pack tests/golden/julia_io   # -> tests/golden/julia_golden.npz (committed)
#
# Not runnable in the build image (no Julia; SURVEY.md F2/F3).  Once tests/golden/julia_golden.npz exists,
# tests/test_oracle.py::test_oracle_matches_julia_reference pins the CPU oracle against the reference itself:
# `F.ipiv` and `F.info` exactly, `F.factors` to the reference's own tolerance (and reports bitwise equality).
# The cases are those of tests/golden/cases.py, which mirror test/runtests.jl:39-64 (sizes 1..10 / 50 / 130 / 300,
# square and s x (s+2), Float64 / Float32, a zeroed column with check = false) plus exact-tie matrices.
#
# index.txt lines:  name m n eltype(f8|f4)
import RecursiveFactorization
using LinearAlgebra

function main(dir)
    for line in eachline(joinpath(dir, "index.txt"))
        isempty(strip(line)) && continue
        name, ms, ns, dt = split(line)
        m, n = parse(Int, ms), parse(Int, ns)
        T = dt == "f4" ? Float32 : Float64
        A = Matrix{T}(undef, m, n)
        read!(joinpath(dir, name * ".in"), A)
        for (tag, thread) in (("serial", Val(false)), ("threaded", Val(true)))
            B = copy(A)
            ipiv = Vector{LinearAlgebra.BlasInt}(undef, min(m, n))
            F = RecursiveFactorization.lu!(B, ipiv, Val(true), thread; check = false)      # src/lu.jl:97-130
            write(joinpath(dir, "$(name).$(tag).factors"), F.factors)
            write(joinpath(dir, "$(name).$(tag).ipiv"), Int64.(F.ipiv))
            write(joinpath(dir, "$(name).$(tag).info"), Int64[F.info])
        end
        B = copy(A)                                                                         # pivot = Val(false), :27-65
        F = RecursiveFactorization.lu!(B, Val(false), Val(false); check = false)
        write(joinpath(dir, "$(name).nopiv.factors"), F.factors)
        write(joinpath(dir, "$(name).nopiv.info"), Int64[F.info])
    end
    open(joinpath(dir, "version.txt"), "w") do io
        println(io, "RecursiveFactorization ", pkgversion(RecursiveFactorization), " julia ", VERSION)
    end
end

main(ARGS[1])
