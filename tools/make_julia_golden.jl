# make_julia_golden.jl — pins the oracle (and the CUDA library) to the REAL reference.
#
# Runs wsshin/StaggeredGridCalculus.jl itself on seeded inputs for every function on the hot path (SURVEY.md §8a) and
# writes inputs AND outputs to tests/golden/julia/{manifest.json,data.bin}.  tests/test_julia_golden.py consumes the
# files when they exist (CPU: oracle == Julia; GPU: libsgc_b200 == Julia — structure bit-exact, values within 4 ulp)
# and reports "parity unpinned" when they do not.  No Julia exists in the build image, so this script has to be run
# by somebody who has one:
#
#     julia --project=/path/to/StaggeredGridCalculus.jl tools/make_julia_golden.jl [outdir]
#
# Only Base, Random, SparseArrays and the package are used (no NPZ / JSON dependency): arrays go to data.bin as raw
# little-endian column-major bytes, the manifest is hand-written JSON.
using StaggeredGridCalculus
using SparseArrays, Random

const OUT = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "julia")
mkpath(OUT)
const BIN = open(joinpath(OUT, "data.bin"), "w")
const CASES = String[]

dtype_name(::Type{Float64}) = "f64"
dtype_name(::Type{ComplexF64}) = "c128"
dtype_name(::Type{Int64}) = "i64"
dtype_name(::Type{Bool}) = "bool"

# one array → raw bytes in data.bin + a JSON descriptor
function arr(A::AbstractArray{T}) where {T}
    B = Array{T}(A)
    off = position(BIN)
    write(BIN, T === Bool ? UInt8.(B) : B)
    return "{\"dtype\":\"$(dtype_name(T))\",\"shape\":[$(join(size(B), ","))],\"offset\":$off}"
end
num(x::Real) = "{\"re\":$(repr(Float64(x))),\"im\":0.0,\"complex\":false}"
num(x::Complex) = "{\"re\":$(repr(Float64(real(x)))),\"im\":$(repr(Float64(imag(x)))),\"complex\":true}"
nums(xs) = "[" * join(num.(xs), ",") * "]"
bools(xs) = "[" * join((x ? "true" : "false" for x in xs), ",") * "]"
ints(xs) = "[" * join(Int.(xs), ",") * "]"
csc(A::SparseMatrixCSC) = "{\"m\":$(size(A,1)),\"n\":$(size(A,2)),\"colptr\":$(arr(Int64.(A.colptr))),\"rowval\":$(arr(Int64.(A.rowval))),\"nzval\":$(arr(A.nzval))}"

function case(fn::String, fields::Pair{String,String}...)
    push!(CASES, "{\"fn\":\"$fn\"," * join(("\"$(k)\":$(v)" for (k, v) in fields), ",") * "}")
end

rng = MersenneTwister(20260)
rfield(T, dims...) = T <: Complex ? rand(rng, dims...) .* 2 .- 1 .+ im .* (rand(rng, dims...) .* 2 .- 1) : rand(rng, dims...) .* 2 .- 1
rcoef(T, n) = T <: Complex ? rand(rng, n) .+ 0.5 .+ im .* (rand(rng, n) .- 0.5) : rand(rng, n) .+ 0.5
ops = (Val(:(=)) => "=", Val(:(+=)) => "+=")

# ---- apply_∂!, apply_m̂!: 1-D, 2-D, 3-D; every axis, direction, boundary, OP; real and complex -------------------------
for N in ((9,), (7, 8), (6, 5, 4), (1, 5, 3)), T in (Float64, ComplexF64)
    K = length(N)
    Fu = rfield(T, N...)
    for nw in 1:K, isfwd in (true, false), isbloch in (true, false), (op, opname) in ops
        isfwd && !isbloch && N[nw] < 2 && continue  # reads Fu[2]
        ∆w⁻¹ = rcoef(T, N[nw]); ∆w = rcoef(T, N[nw]); ∆w′⁻¹ = rcoef(T, N[nw])
        e = T <: Complex ? exp(-im * 0.7) : -1.0
        α = T <: Complex ? 0.5 - 0.25im : 1.5
        G0 = rfield(T, N...)
        G = copy(G0); apply_∂!(G, Fu, op, nw, isfwd, ∆w⁻¹, isbloch, e; α=α)
        case("apply_partial", "N" => ints(N), "op" => "\"$opname\"", "nw" => "$nw", "isfwd" => "$isfwd", "isbloch" => "$isbloch",
             "e" => num(e), "alpha" => num(α), "dwinv" => arr(∆w⁻¹), "F" => arr(Fu), "G0" => arr(G0), "G" => arr(G))
        G = copy(G0); apply_m̂!(G, Fu, op, nw, isfwd, isbloch, e; α=α)
        case("apply_mhat", "N" => ints(N), "op" => "\"$opname\"", "nw" => "$nw", "isfwd" => "$isfwd", "isbloch" => "$isbloch",
             "e" => num(e), "alpha" => num(α), "F" => arr(Fu), "G0" => arr(G0), "G" => arr(G))
        G = copy(G0); apply_m̂!(G, Fu, op, nw, isfwd, ∆w, ∆w′⁻¹, isbloch, e; α=α)
        case("apply_mhat_weighted", "N" => ints(N), "op" => "\"$opname\"", "nw" => "$nw", "isfwd" => "$isfwd", "isbloch" => "$isbloch",
             "e" => num(e), "alpha" => num(α), "dw" => arr(∆w), "dwpinv" => arr(∆w′⁻¹), "F" => arr(Fu), "G0" => arr(G0), "G" => arr(G))
    end
end

# ---- apply_curl! (all 8 direction triples, mixed boundaries incl. backward Bloch with complex e), divg, grad, mean ----
let N = (8, 7, 6), T = ComplexF64
    F = rfield(T, N..., 3); f = rfield(T, N...)
    ∆l⁻¹ = ntuple(d -> rcoef(T, N[d]), 3); ∆l = ntuple(d -> rcoef(T, N[d]), 3)
    e = [exp(-im * 0.3), exp(im * 0.5), exp(-im * 0.9)]
    for fx in (true, false), fy in (true, false), fz in (true, false), isbloch in ([true, true, true], [true, false, true], [false, false, false])
        isfwd = [fx, fy, fz]
        for (op, opname) in ops
            G0 = rfield(T, N..., 3); g0 = rfield(T, N...)
            G = copy(G0); apply_curl!(G, F, op, isfwd, ∆l⁻¹, isbloch, e; α=0.5 - 0.25im)
            case("apply_curl", "N" => ints(N), "op" => "\"$opname\"", "isfwd" => bools(isfwd), "isbloch" => bools(isbloch), "e" => nums(e),
                 "alpha" => num(0.5 - 0.25im), "dlinv" => "[" * join(arr.(∆l⁻¹), ",") * "]", "F" => arr(F), "G0" => arr(G0), "G" => arr(G))
            g = copy(g0); apply_divg!(g, F, op, isfwd, ∆l⁻¹, isbloch, e)
            case("apply_divg", "N" => ints(N), "op" => "\"$opname\"", "isfwd" => bools(isfwd), "isbloch" => bools(isbloch), "e" => nums(e),
                 "dlinv" => "[" * join(arr.(∆l⁻¹), ",") * "]", "F" => arr(F), "G0" => arr(g0), "G" => arr(g))
            G = copy(G0); apply_grad!(G, f, op, isfwd, ∆l⁻¹, isbloch, e)
            case("apply_grad", "N" => ints(N), "op" => "\"$opname\"", "isfwd" => bools(isfwd), "isbloch" => bools(isbloch), "e" => nums(e),
                 "dlinv" => "[" * join(arr.(∆l⁻¹), ",") * "]", "F" => arr(f), "G0" => arr(G0), "G" => arr(G))
            G = copy(G0); apply_mean!(G, F, op, isfwd, ∆l, ∆l⁻¹, isbloch, e)
            case("apply_mean_weighted", "N" => ints(N), "op" => "\"$opname\"", "isfwd" => bools(isfwd), "isbloch" => bools(isbloch), "e" => nums(e),
                 "dl" => "[" * join(arr.(∆l), ",") * "]", "dlpinv" => "[" * join(arr.(∆l⁻¹), ",") * "]", "F" => arr(F), "G0" => arr(G0), "G" => arr(G))
        end
    end
    # 2-D TE / TM sub-curls (test/curl.jl:67-92)
    N2 = (9, 8); d2 = ntuple(d -> rcoef(Float64, N2[d]), 2)
    F2 = rfield(Float64, N2..., 2); F1 = rfield(Float64, N2..., 1)
    for isfwd in ([true, true], [false, true]), isbloch in ([true, true], [false, false])
        G = zeros(N2..., 1); apply_curl!(G, F2, Val(:(=)), isfwd, d2, isbloch, ones(2); cmp_shp=[1, 2], cmp_out=[3], cmp_in=[1, 2])
        case("apply_curl_sub", "N" => ints(N2), "isfwd" => bools(isfwd), "isbloch" => bools(isbloch), "cmp_out" => ints([3]), "cmp_in" => ints([1, 2]),
             "dlinv" => "[" * join(arr.(d2), ",") * "]", "F" => arr(F2), "G" => arr(G))
        G = zeros(N2..., 2); apply_curl!(G, F1, Val(:(=)), isfwd, d2, isbloch, ones(2); cmp_shp=[1, 2], cmp_out=[1, 2], cmp_in=[3])
        case("apply_curl_sub", "N" => ints(N2), "isfwd" => bools(isfwd), "isbloch" => bools(isbloch), "cmp_out" => ints([1, 2]), "cmp_in" => ints([3]),
             "dlinv" => "[" * join(arr.(d2), ",") * "]", "F" => arr(F1), "G" => arr(G))
    end
end

# ---- sparse assembly: every create_*, both orderings, real and complex, Bloch and symmetric ---------------------------
for N in ((6,), (5, 4), (4, 3, 5), (1, 4, 2)), T in (Float64, ComplexF64)
    K = length(N)
    ∆l⁻¹ = ntuple(d -> rcoef(T, N[d]), K); ∆l = ntuple(d -> rcoef(T, N[d]), K)
    e = T <: Complex ? [exp(-im * 0.3), exp(im * 0.5), exp(-im * 0.9)][1:K] : ones(K)
    for isbloch in (fill(true, K), fill(false, K)), isfwd in (fill(true, K), fill(false, K), [isodd(d) for d in 1:K])
        for nw in 1:K
            A = create_∂(nw, isfwd[nw], [N...], ∆l⁻¹[nw], isbloch[nw], e[nw])
            case("create_partial", "N" => ints(N), "nw" => "$nw", "isfwd" => "$(isfwd[nw])", "isbloch" => "$(isbloch[nw])", "e" => num(e[nw]),
                 "dwinv" => arr(∆l⁻¹[nw]), "A" => csc(A))
            A = create_m̂(nw, isfwd[nw], [N...], ∆l[nw], ∆l⁻¹[nw], isbloch[nw], e[nw])
            case("create_mhat", "N" => ints(N), "nw" => "$nw", "isfwd" => "$(isfwd[nw])", "isbloch" => "$(isbloch[nw])", "e" => num(e[nw]),
                 "dw" => arr(∆l[nw]), "dwpinv" => arr(∆l⁻¹[nw]), "A" => csc(A))
        end
        for order_cmpfirst in (true, false)
            common = ("N" => ints(N), "isfwd" => bools(isfwd), "isbloch" => bools(isbloch), "e" => nums(e), "order_cmpfirst" => "$order_cmpfirst",
                      "dlinv" => "[" * join(arr.(∆l⁻¹), ",") * "]")
            K == 3 && case("create_curl", common..., "A" => csc(create_curl(isfwd, ∆l⁻¹, isbloch, e; order_cmpfirst)))
            case("create_divg", common..., "A" => csc(create_divg(isfwd, ∆l⁻¹, isbloch, e; order_cmpfirst)))
            case("create_grad", common..., "A" => csc(create_grad(isfwd, ∆l⁻¹, isbloch, e; order_cmpfirst)))
            case("create_mean", common..., "dl" => "[" * join(arr.(∆l), ",") * "]",
                 "A" => csc(create_mean(isfwd, ∆l, ∆l⁻¹, isbloch, e; order_cmpfirst)))
        end
    end
    K == 3 && for order_cmpfirst in (true, false)
        sc = rcoef(T, 3)
        case("create_picmp", "N" => ints(N), "permute" => ints([2, 3, 1]), "scale" => arr(sc), "order_cmpfirst" => "$order_cmpfirst",
             "A" => csc(create_πcmp([N...], [2, 3, 1]; scale=sc, order_cmpfirst)))
    end
end

# ---- grid / PML producers of the per-axis vectors (src/grid.jl, src/pml.jl, src/util.jl) -------------------------------
let lprim = ([0.0; cumsum(rand(rng, 12) .+ 0.5)], [0.0; cumsum(rand(rng, 10) .+ 0.5)], [0.0; cumsum(rand(rng, 9) .+ 0.5)])
    for isbloch in ([false, false, false], [true, false, true])
        g = Grid(lprim, isbloch)
        ω = 2π / 7.3
        Npml = ([3, 2, 0], [2, 3, 1])
        s∆l = create_stretched_∆l(ω, g, Npml)
        inv∆l = invert_∆l(s∆l)
        case("pml", "isbloch" => bools(isbloch), "omega" => repr(ω), "npml_neg" => ints(Npml[1]), "npml_pos" => ints(Npml[2]),
             "lprim" => "[" * join(arr.(lprim), ",") * "]",
             "sdl_prim" => "[" * join(arr.(s∆l[1]), ",") * "]", "sdl_dual" => "[" * join(arr.(s∆l[2]), ",") * "]",
             "inv_prim" => "[" * join(arr.(inv∆l[1]), ",") * "]", "inv_dual" => "[" * join(arr.(inv∆l[2]), ",") * "]")
    end
end

close(BIN)
open(joinpath(OUT, "manifest.json"), "w") do io
    write(io, "{\"generator\":\"tools/make_julia_golden.jl\",\"julia\":\"$(VERSION)\",\"cases\":[\n" * join(CASES, ",\n") * "\n]}\n")
end
println("wrote $(length(CASES)) cases to $OUT")
