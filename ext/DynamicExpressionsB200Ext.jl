# DynamicExpressionsB200Ext — package extension binding libdexb200.so (include/dexb200.h) behind the
# evaluation API of DynamicExpressions.jl.
#
# Pattern: the reference's own weak-dependency extensions (ext/DynamicExpressionsBumperExt.jl:11-49
# implements `_bumper_eval_tree_array`, dispatched from src/Evaluate.jl:300-302 through the stubs of
# src/ExtensionInterface.jl:39-64).  This one adds METHODS of the public entry points for
#   * a device-matrix wrapper `B200Matrix{T}` (single tree: drop-in for `eval_tree_array(tree, X, ops)`),
#   * vectors of trees (the batched form the reference lacks; callers write
#     `[eval_tree_array(t, X, ops) for t in trees]`, benchmark/benchmarks.jl:76-91),
# for eval / eval_grad / eval_diff / ParametricExpression / fused loss, all through `ccall`.
#
# STATUS: Julia is not installed in the build environment of this repository, so this file has
# never been executed.  tests/c_abi/julia_ext_replay.c replays, from plain C, the exact sequence of
# library calls made below (same argument order, same `DexNode` layout, 0-based conversion, column-
# major result interpretation) and is run on the GPU by the test-suite; a maintainer with Julia adds
# this file as `ext/DynamicExpressionsB200Ext.jl` plus a `[weakdeps]`/`[extensions]` entry keyed on a
# tiny `DynamicExpressionsB200` trigger package that only carries the path of the shared library.
module DynamicExpressionsB200Ext

using DynamicExpressions:
    AbstractExpressionNode,
    AbstractExpression,
    OperatorEnum,
    EvalContext,
    ParametricExpression,
    ParametricNode,
    count_nodes,
    get_child,
    get_tree,
    get_operators,
    get_metadata
import DynamicExpressions: eval_tree_array, eval_grad_tree_array, eval_diff_tree_array

const LIB = get(ENV, "DEXB200_LIB", "libdexb200.so")

# ---- include/dex_wire.h ---------------------------------------------------------------------
# image of `struct dex_node`, 16 bytes: offsets 0 degree, 1 kind, 2 op, 3 pad, 4 feature, 6 pad, 8 val
struct DexNode
    degree::UInt8
    kind::UInt8      # leaves: 0 constant, 1 feature, 2 parameter
    op::UInt8        # 0-based index into operators[degree]
    reserved0::UInt8
    feature::UInt16  # 0-based feature / parameter row
    reserved1::UInt16
    val::Float64
end
@assert sizeof(DexNode) == 16 && fieldoffset(DexNode, 5) == 4 && fieldoffset(DexNode, 7) == 8

const DEX_F32, DEX_F64 = Cint(0), Cint(1)
const DEX_EVAL_EARLY_EXIT = Cint(1)
const DEX_EVAL_SKIP_INCOMPLETE = Cint(2)   # dex_eval_host: rows of incomplete trees are not transferred
const DEX_PACK_FUSED, DEX_PACK_BUMPER = Cint(1), Cint(2)
const DEX_GRAD_CONSTANTS, DEX_GRAD_FEATURES, DEX_GRAD_BOTH = Cint(0), Cint(1), Cint(2)
dtype_code(::Type{Float32}) = DEX_F32
dtype_code(::Type{Float64}) = DEX_F64
dtype_code(::Type{T}) where {T} = error("the B200 path evaluates Float32 / Float64 only, got $T")

"""Host matrix whose evaluation is routed to the GPU: `eval_tree_array(tree, B200Matrix(X), operators)`."""
struct B200Matrix{T,M<:AbstractMatrix{T}} <: AbstractMatrix{T}
    data::M
    device::Int
end
B200Matrix(X::AbstractMatrix; device::Integer=0) = B200Matrix{eltype(X),typeof(X)}(X, Int(device))
Base.size(X::B200Matrix) = size(X.data)
Base.getindex(X::B200Matrix, i::Int, j::Int) = X.data[i, j]

# ---- contexts: one per (task, device), like the reference's task-local Bumper slab ----------------
function context(device::Integer=0)
    get!(task_local_storage(), (:dexb200_ctx, Int(device))) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:dex_ctx_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h)
        rc == 0 || error("dex_ctx_create: ", unsafe_string(ccall((:dex_strerror, LIB), Cstring, (Cint,), rc)))
        h[]
    end::Ptr{Cvoid}
end
function check(ctx, rc)
    rc == 0 || error(unsafe_string(ccall((:dex_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))
    return nothing
end

# ---- OperatorEnum -> builtin opcodes (cached per OperatorEnum object) ------------------------------
# `declare_operator_alias`-style names of the reference's test-suite (safe_log, ...) resolve through
# the alias column of include/dex_ops.def; a function without a device implementation is an error at
# table-build time: there is no CPU fallback.
const OPTABLES = IdDict{Any,Ptr{Cvoid}}()
const OPTABLE_LOCK = ReentrantLock()
function optable(operators::OperatorEnum)
    lock(OPTABLE_LOCK) do
        get!(OPTABLES, operators) do
            codes = Int32[]
            offs = Int32[0]
            for (d, ops) in enumerate(operators.ops)
                for f in ops
                    c = ccall((:dex_opcode_from_name, LIB), Cint, (Cstring, Cint), string(nameof(f)), d)
                    c >= 0 || error("operator `$(nameof(f))` of degree $d has no B200 implementation (include/dex_ops.def)")
                    push!(codes, c)
                end
                push!(offs, length(codes))
            end
            h = Ref{Ptr{Cvoid}}(C_NULL)
            rc = ccall((:dex_optable_create, LIB), Cint, (Ptr{Int32}, Ptr{Int32}, Cint, Ref{Ptr{Cvoid}}),
                       codes, offs, length(operators.ops), h)
            rc == 0 || error("dex_optable_create: ", unsafe_string(ccall((:dex_strerror, LIB), Cstring, (Cint,), rc)))
            h[]
        end
    end
end

# ---- flattening: parent first, children left to right (the order of tree_mapreduce, src/base.jl:123-158,
# which is also the constant numbering of index_constant_nodes, src/NodeUtils.jl:184-201) -------------
function flatten!(out::Vector{DexNode}, tree::AbstractExpressionNode)
    d = tree.degree
    if d == 0
        if tree.constant
            push!(out, DexNode(0, 0, 0, 0, 0, 0, Float64(tree.val)))
        elseif tree isa ParametricNode && tree.is_parameter
            push!(out, DexNode(0, 2, 0, 0, UInt16(tree.parameter - 1), 0, 0.0))
        else
            push!(out, DexNode(0, 1, 0, 0, UInt16(tree.feature - 1), 0, 0.0))
        end
    else
        push!(out, DexNode(d, 0, UInt8(tree.op - 1), 0, 0, 0, 0.0))
        for i in 1:d
            flatten!(out, get_child(tree, i))
        end
    end
    return out
end

# ---- packed populations, cached by the identity of the trees and their structure hash -----------------
mutable struct Population
    handle::Ptr{Cvoid}
    ctx::Ptr{Cvoid}
    n_trees::Int
    n_constants::Vector{Int32}
end
destroy!(p::Population) = (p.handle == C_NULL || ccall((:dex_population_destroy, LIB), Cint, (Ptr{Cvoid},), p.handle); p.handle = C_NULL; nothing)

const POPULATIONS = Dict{Tuple{UInt,Ptr{Cvoid},Cint,Cint},Population}()   # LRU of a few entries is enough
function population(trees::AbstractVector{<:AbstractExpressionNode{T}}, operators::OperatorEnum, ctx;
                    pack_flags::Cint=DEX_PACK_FUSED, n_params::Integer=0) where {T}
    nodes = DexNode[]
    offsets = Int64[0]
    for t in trees
        flatten!(nodes, t)
        push!(offsets, length(nodes))
    end
    flags = pack_flags | Cint((n_params & 0xffff) << 8)                   # DEX_PACK_PARAM_ROWS(n_params)
    key = (hash(nodes, hash(offsets)), ctx, dtype_code(T), flags)
    p = get(POPULATIONS, key, nothing)
    p === nothing || return p
    length(POPULATIONS) > 16 && (foreach(destroy!, values(POPULATIONS)); empty!(POPULATIONS))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:dex_population_pack, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{DexNode}, Ptr{Int64}, Int64, Cint, Cint, Ref{Ptr{Cvoid}}),
                     ctx, optable(operators), nodes, offsets, length(trees), dtype_code(T), flags, h))
    counts = Vector{Int32}(undef, length(trees))
    ccall((:dex_population_constant_counts, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), h[], counts)
    p = Population(h[], ctx, length(trees), counts)
    finalizer(destroy!, p)
    POPULATIONS[key] = p
    return p
end

pack_flags(ec::Union{EvalContext,Nothing}) =
    ec === nothing ? DEX_PACK_FUSED :
    ((ec.use_fused isa Val{true} ? DEX_PACK_FUSED : Cint(0)) | (ec.bumper isa Val{true} ? DEX_PACK_BUMPER : Cint(0)))
# with early exit the reference hands back an unusable buffer for an incomplete tree
# (src/Evaluate.jl:26-32); the library then does not even transfer that column
eval_flags(ec::Union{EvalContext,Nothing}) =
    (ec === nothing || ec.early_exit isa Val{true}) ? (DEX_EVAL_EARLY_EXIT | DEX_EVAL_SKIP_INCOMPLETE) : Cint(0)

# ---- device buffers as raw pointers (dex_device_alloc / dex_copy_to_device / dex_copy_to_host) ------
function with_device(f, ctx, bytes::Integer...)
    ptrs = Ptr{Cvoid}[]
    try
        for b in bytes
            r = Ref{Ptr{Cvoid}}(C_NULL)
            check(ctx, ccall((:dex_device_alloc, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Int64), ctx, r, b))
            push!(ptrs, r[])
        end
        return f(ptrs...)
    finally
        foreach(p -> ccall((:dex_device_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx, p), ptrs)
    end
end
to_device(ctx, dst, src::Array) = check(ctx, ccall((:dex_copy_to_device, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), ctx, dst, src, sizeof(src)))
to_host(ctx, dst::Array, src) = check(ctx, ccall((:dex_copy_to_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), ctx, dst, src, sizeof(dst)))

# =========================================================================================
# eval_tree_array
# =========================================================================================
"""Batched `[eval_tree_array(t, X, operators) for t in trees]` in one launch.
Returns `(out::Matrix{T} (nsamples x ntrees), complete::Vector{Bool})`: column `t` is the vector the
reference returns for tree `t` (the library's row-major (ntrees x nsamples) result IS this matrix)."""
function eval_tree_array(trees::AbstractVector{<:AbstractExpressionNode{T}}, X::B200Matrix{T}, operators::OperatorEnum;
                         eval_context::Union{EvalContext,Nothing}=nothing) where {T}
    ctx = context(X.device)
    pop = population(trees, operators, ctx; pack_flags=pack_flags(eval_context))
    cX = X.data isa Matrix{T} ? X.data : Matrix{T}(X.data)
    F, N = size(cX)
    out = Matrix{T}(undef, N, length(trees))
    ok = Vector{UInt8}(undef, length(trees))
    check(ctx, ccall((:dex_eval_host, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{T}, Int32, Int64, Int64, Ptr{T}, Int64, Ptr{UInt8}, Cint),
                     ctx, pop.handle, cX, F, N, F, out, N, ok, eval_flags(eval_context)))
    return out, ok .!= 0
end
"""The same over several devices of this process: device `devices[d]` evaluates the d-th contiguous
column block of `X` against its own packed copy of the trees and its rows land in place in the result
(`dex_shard_eval_host`: every device is enqueued before any is waited for)."""
function eval_tree_array(trees::AbstractVector{<:AbstractExpressionNode{T}}, X::B200Matrix{T}, operators::OperatorEnum,
                         devices::AbstractVector{<:Integer}; eval_context::Union{EvalContext,Nothing}=nothing) where {T}
    ctxs = Ptr{Cvoid}[context(d) for d in devices]
    pops = [population(trees, operators, c; pack_flags=pack_flags(eval_context)) for c in ctxs]
    handles = Ptr{Cvoid}[p.handle for p in pops]
    cX = X.data isa Matrix{T} ? X.data : Matrix{T}(X.data)
    F, N = size(cX)
    out = Matrix{T}(undef, N, length(trees))
    ok = Vector{UInt8}(undef, length(trees))
    GC.@preserve pops check(ctxs[1], ccall((:dex_shard_eval_host, LIB), Cint,
                     (Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Int32, Ptr{T}, Int32, Int64, Int64, Ptr{T}, Int64, Ptr{UInt8}, Cint),
                     ctxs, handles, length(devices), cX, F, N, F, out, N, ok, eval_flags(eval_context)))
    return out, ok .!= 0
end
# the reference's single-tree signature (src/Evaluate.jl:279-285)
function eval_tree_array(tree::AbstractExpressionNode{T}, X::B200Matrix{T}, operators::OperatorEnum;
                         eval_context::Union{EvalContext,Nothing}=nothing, kws...) where {T}
    out, ok = eval_tree_array([tree], X, operators; eval_context)
    return vec(out), ok[1]
end

# =========================================================================================
# eval_grad_tree_array / eval_diff_tree_array (src/EvaluateDerivative.jl:40-53, 193-228)
# =========================================================================================
grad_mode(::Val{true}) = DEX_GRAD_FEATURES
grad_mode(::Val{false}) = DEX_GRAD_CONSTANTS
grad_mode(::Val{:both}) = DEX_GRAD_BOTH
grad_mode(b::Bool) = b ? DEX_GRAD_FEATURES : DEX_GRAD_CONSTANTS

"""Returns `(evaluation::Matrix (N x P), gradients::Vector{Matrix} (G_t x N each), complete::Vector{Bool})`."""
function eval_grad_tree_array(trees::AbstractVector{<:AbstractExpressionNode{T}}, X::B200Matrix{T}, operators::OperatorEnum;
                              variable::Union{Bool,Val}=Val(false), kws...) where {T}
    ctx = context(X.device)
    pop = population(trees, operators, ctx)
    cX = X.data isa Matrix{T} ? X.data : Matrix{T}(X.data)
    F, N = size(cX)
    P = length(trees)
    mode = grad_mode(variable)
    offs = Vector{Int64}(undef, P + 1)
    ccall((:dex_grad_offsets, LIB), Cint, (Ptr{Cvoid}, Int32, Int64, Cint, Ptr{Int64}), pop.handle, F, N, mode, offs)
    out = Matrix{T}(undef, N, P)
    grad = Vector{T}(undef, offs[end])
    ok = Vector{UInt8}(undef, P)
    with_device(ctx, sizeof(cX), sizeof(out), max(sizeof(grad), 1), P) do dX, dO, dG, dK
        to_device(ctx, dX, cX)
        check(ctx, ccall((:dex_eval_grad, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Cint, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Int64}, Ptr{Cvoid}),
                         ctx, pop.handle, dX, F, N, F, mode, dO, N, dG, offs, dK))
        to_host(ctx, out, dO); to_host(ctx, grad, dG); to_host(ctx, ok, dK)
    end
    # tree t's block is (G_t x N) column-major, gradient index fastest: exactly Julia's Matrix memory
    grads = [reshape(view(grad, (offs[t] + 1):offs[t + 1]), :, N) for t in 1:P]
    return out, grads, ok .!= 0
end
function eval_grad_tree_array(tree::AbstractExpressionNode{T}, X::B200Matrix{T}, operators::OperatorEnum; kws...) where {T}
    out, grads, ok = eval_grad_tree_array([tree], X, operators; kws...)
    return vec(out), Matrix(grads[1]), ok[1]
end
function eval_diff_tree_array(tree::AbstractExpressionNode{T}, X::B200Matrix{T}, operators::OperatorEnum, direction::Integer; kws...) where {T}
    ctx = context(X.device)
    pop = population([tree], operators, ctx)
    cX = X.data isa Matrix{T} ? X.data : Matrix{T}(X.data)
    F, N = size(cX)
    out = Vector{T}(undef, N); dout = Vector{T}(undef, N); ok = Vector{UInt8}(undef, 1)
    with_device(ctx, sizeof(cX), sizeof(out), sizeof(dout), 1) do dX, dO, dD, dK
        to_device(ctx, dX, cX)
        check(ctx, ccall((:dex_eval_diff, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
                         ctx, pop.handle, dX, F, N, F, direction - 1, dO, dD, N, dK))
        to_host(ctx, out, dO); to_host(ctx, dout, dD); to_host(ctx, ok, dK)
    end
    return out, dout, ok[1] != 0
end

# =========================================================================================
# ParametricExpression (src/ParametricExpression.jl:371-390): the gather parameters[:, classes] and the
# vcat onto X happen inside the kernel's operand fetch
# =========================================================================================
function eval_tree_array(exs::AbstractVector{<:ParametricExpression{T}}, X::B200Matrix{T}, classes::AbstractVector{<:Integer},
                         operators::Union{OperatorEnum,Nothing}=nothing; eval_context::Union{EvalContext,Nothing}=nothing) where {T}
    ctx = context(X.device)
    ops = get_operators(first(exs), operators)
    n_params, n_classes = size(get_metadata(first(exs)).parameters)
    @assert length(classes) == size(X, 2) && maximum(classes) <= n_classes            # :378-379
    pop = population([get_tree(ex) for ex in exs], ops, ctx; pack_flags=pack_flags(eval_context), n_params)
    cX = X.data isa Matrix{T} ? X.data : Matrix{T}(X.data)
    F, N = size(cX)
    P = length(exs)
    params = Array{T,3}(undef, n_params, n_classes, P)                                # per tree column-major
    for (t, ex) in enumerate(exs)
        params[:, :, t] .= get_metadata(ex).parameters
    end
    cls0 = Int32.(classes .- 1)
    out = Matrix{T}(undef, N, P); ok = Vector{UInt8}(undef, P)
    with_device(ctx, sizeof(cX), max(sizeof(params), 1), sizeof(cls0), sizeof(out), P) do dX, dP, dC, dO, dK
        to_device(ctx, dX, cX); to_device(ctx, dP, params); to_device(ctx, dC, cls0)
        check(ctx, ccall((:dex_eval_parametric, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Cint),
                         ctx, pop.handle, dX, F, N, F, dP, n_params, n_classes, dC, dO, N, dK, eval_flags(eval_context)))
        to_host(ctx, out, dO); to_host(ctx, ok, dK)
    end
    return out, ok .!= 0
end
function eval_tree_array(ex::ParametricExpression{T}, X::B200Matrix{T}, classes::AbstractVector{<:Integer},
                         operators::Union{OperatorEnum,Nothing}=nothing; kws...) where {T}
    out, ok = eval_tree_array([ex], X, classes, operators; kws...)
    return vec(out), ok[1]
end

# =========================================================================================
# fused loss (what optimiser callers consume, test/test_optim.jl:44-52; the pullback's contraction,
# src/ChainRules.jl:56-77): neither the (P x N) values nor the (G x N) gradients leave the GPU
# =========================================================================================
"""`loss[t] = sum_j w_j (tree_t(X[:, j]) - y[j])^2 / sum_j w_j` and, with `variable`, its gradient
w.r.t. constants / features: `(loss::Vector{Float64}, grads::Vector{Vector{Float64}}, complete)`."""
function eval_loss_and_grad(trees::AbstractVector{<:AbstractExpressionNode{T}}, X::B200Matrix{T}, y::AbstractVector{T},
                            operators::OperatorEnum; weights::Union{Nothing,AbstractVector{T}}=nothing,
                            variable::Union{Bool,Val}=Val(false)) where {T}
    ctx = context(X.device)
    pop = population(trees, operators, ctx)
    cX = X.data isa Matrix{T} ? X.data : Matrix{T}(X.data)
    F, N = size(cX)
    P = length(trees)
    mode = grad_mode(variable)
    offs = Vector{Int64}(undef, P + 1)
    ccall((:dex_grad_offsets, LIB), Cint, (Ptr{Cvoid}, Int32, Int64, Cint, Ptr{Int64}), pop.handle, F, 1, mode, offs)
    loss = Vector{Float64}(undef, P); grad = Vector{Float64}(undef, offs[end]); ok = Vector{UInt8}(undef, P)
    yv = Vector{T}(y); wv = weights === nothing ? T[] : Vector{T}(weights)
    with_device(ctx, sizeof(cX), sizeof(yv), max(sizeof(wv), 1), sizeof(loss), max(sizeof(grad), 1), P) do dX, dY, dW, dL, dG, dK
        to_device(ctx, dX, cX); to_device(ctx, dY, yv)
        weights === nothing || to_device(ctx, dW, wv)
        check(ctx, ccall((:dex_eval_loss_grad, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Cvoid}),
                         ctx, pop.handle, dX, F, N, F, dY, weights === nothing ? C_NULL : dW, mode, dL, dG, offs, dK))
        to_host(ctx, loss, dL); to_host(ctx, grad, dG); to_host(ctx, ok, dK)
    end
    return loss, [grad[(offs[t] + 1):offs[t + 1]] for t in 1:P], ok .!= 0
end

# get/set_scalar_constants for a packed population without re-packing (src/NodeUtils.jl:99-143;
# serves ext/DynamicExpressionsOptimExt.jl:192-224: constants change every BFGS iteration)
function set_constants!(pop::Population, values::Vector{T}) where {T<:Union{Float32,Float64}}
    check(pop.ctx, ccall((:dex_population_set_constants, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{T}, Int64),
                         pop.ctx, pop.handle, values, length(values)))
end

end # module
