# Times the REAL reference on this box's host cores (used by `bench.py --impl reference` when a
# `julia` binary with DynamicExpressions.jl is available; this container and the GPU boxes have
# none, so this script has never been executed here — bench.py falls back to the C port).
#
#   julia --threads=N benchmarks/reference_julia.jl <dir with nodes.npy offsets.npy X.npy> <steps> <warmup>
#
# It rebuilds the population from the wire dump (include/dex_wire.h: 16-byte preorder records),
# then times the loop of /root/reference/benchmark/benchmarks.jl:76-91,
#   [eval_tree_array(tree, X, operators; eval_context...) for tree in trees],
# threaded over trees (how SymbolicRegression.jl drives it), for the plain, turbo and bumper
# variants, and prints one JSON line.
using DynamicExpressions
using DynamicExpressions: EvalOptions
try
    @eval using LoopVectorization
catch
end
try
    @eval using Bumper
catch
end

function read_npy(path)
    io = open(path)
    magic = read(io, 6); ver = read(io, 2)
    hlen = ver[1] == 1 ? Int(read(io, UInt16)) : Int(read(io, UInt32))
    header = String(read(io, hlen))
    descr = match(r"'descr': '([^']+)'", header).captures[1]
    shape = [parse(Int, m.match) for m in eachmatch(r"\d+", match(r"'shape': \(([^)]*)\)", header).captures[1])]
    fortran = occursin("'fortran_order': True", header)
    T = Dict("|u1" => UInt8, "<i8" => Int64, "<f4" => Float32, "<f8" => Float64)[descr]
    data = Vector{T}(undef, prod(shape)); read!(io, data); close(io)
    return data, shape, fortran
end

const OPS = OperatorEnum(; binary_operators=[+, -, /, *], unary_operators=[cos, exp])   # benchmark/benchmarks.jl:32-36

# one preorder record: degree u8, kind u8, op u8, pad u8, feature u16, pad u16, val f64 (0-based indices)
function build(bytes::Vector{UInt8}, pos::Ref{Int})
    base = pos[] * 16
    degree, kind, op = bytes[base + 1], bytes[base + 2], bytes[base + 3]
    feature = reinterpret(UInt16, bytes[(base + 5):(base + 6)])[1]
    val = reinterpret(Float64, bytes[(base + 9):(base + 16)])[1]
    pos[] += 1
    if degree == 0
        return kind == 0 ? Node(Float32; val=Float32(val)) : Node(Float32; feature=Int(feature) + 1)
    elseif degree == 1
        l = build(bytes, pos)
        return Node(Int(op) + 1, l)
    else
        l = build(bytes, pos); r = build(bytes, pos)
        return Node(Int(op) + 1, l, r)
    end
end

function main()
    dir, steps, warmup = ARGS[1], parse(Int, ARGS[2]), parse(Int, ARGS[3])
    bytes, _, _ = read_npy(joinpath(dir, "nodes.npy"))
    offsets, _, _ = read_npy(joinpath(dir, "offsets.npy"))
    xdata, xshape, _ = read_npy(joinpath(dir, "X.npy"))            # (F, N) C-order
    X = permutedims(reshape(xdata, xshape[2], xshape[1]))           # -> F x N column-major Matrix{Float32}
    trees = [build(bytes, Ref(Int(offsets[t]))) for t in 1:(length(offsets) - 1)]
    nodeops = sum(count_nodes, trees) * size(X, 2)
    results = Dict{String,Any}()
    variants = [("plain", (turbo=false, bumper=false))]
    isdefined(Main, :LoopVectorization) && push!(variants, ("turbo", (turbo=true, bumper=false)))
    isdefined(Main, :Bumper) && push!(variants, ("bumper", (turbo=false, bumper=true)))
    (isdefined(Main, :LoopVectorization) && isdefined(Main, :Bumper)) && push!(variants, ("turbo+bumper", (turbo=true, bumper=true)))
    best = Inf
    for (name, kw) in variants
        opts = EvalOptions(; kw...)
        step() = Threads.@threads for i in eachindex(trees)
            eval_tree_array(trees[i], X, OPS; eval_options=opts)
        end
        for _ in 1:warmup; step(); end
        t = @elapsed for _ in 1:steps; step(); end
        results[name] = t / steps * 1e3
        best = min(best, t / steps)
    end
    msg = "DynamicExpressions.jl eval_tree_array over the population, Threads.@threads over trees on $(Threads.nthreads()) threads; best of $(join(first.(variants), ", "))"
    print("{\"ms_per_step\": $(best * 1e3), \"node_ops_per_s\": $(nodeops / best), \"threads\": $(Threads.nthreads()), ")
    print("\"variants_ms\": {", join(["\"$k\": $v" for (k, v) in results], ", "), "}, \"sample\": \"$msg\"}\n")
end

main()
