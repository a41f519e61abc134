# GPUSink.jl — the reference-side binding for libsignalops_cuda.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  This file is the
# glue a SignalOperators.jl maintainer adds (e.g. as src/GPUSink.jl behind a `Requires`
# hook next to src/SignalOperators.jl:45-72).  It mirrors, 1:1, the Python host layer
# that IS exercised here (signaloperators.jl_b200/host/lowering.py emits the same plan
# bytes; tests/plan_emulator.py documents their meaning).
#
# Plug-in point: the documented custom-sink interface (docs/src/custom_sink.md:1-18):
# `sink(x, to)` dispatches on `to`, exactly like `sink(x, to::String)` at
# src/sink.jl:139-142.  Nothing of the block-pull machinery (`nextblock`, `frame`,
# `sink_helper!`) is used.

module GPUSinks

using SignalOperators
using SignalOperators: AbstractSignal, CutApply, PaddedSignal, AppendSignals, FilteredSignal,
    NormedSignal, MapSignal, RampSignal, SignalFunction, NumberSignal, FilterFn, RawFilterFn,
    ResamplerFn, FnBr, ToEltypeFn, AsNChannels, As1Channel, GetChanFn, tuplecat, RandFn,
    process_sink_params, initsink, refineroot, root, resolvelen, child, sinramp, inflen
using DSP

export GPUSink

const libsignalops = "libsignalops_cuda"

# ---- C ABI (include/signalops.h) -------------------------------------------------------
struct SigopsBuffer
    ptr::Ptr{Cvoid}
    nframes::Int64
    nchannels::Int32
    dtype::Int32          # 1 = Float32, 2 = Float64, 3 = Int64
    ld::Int64
end

mutable struct SigopsStats
    gpu_ms::Float64; h2d_ms::Float64; d2h_ms::Float64; wall_ms::Float64
    launches::Int64; h2d_bytes::Int64; d2h_bytes::Int64; out_samples::Int64
    SigopsStats() = new(0, 0, 0, 0, 0, 0, 0, 0)
end

mutable struct GPUSink
    devices::Vector{Cint}
    ctx::Ptr{Cvoid}
    plans::Dict{Vector{UInt8},Ptr{Cvoid}}
    function GPUSink(devices = [0])
        ctx = Ref{Ptr{Cvoid}}(C_NULL)
        devs = Cint.(devices)
        rc = ccall((:sigops_ctx_create, libsignalops), Cint, (Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}),
                   devs, length(devs), ctx)
        rc == 0 || error(unsafe_string(ccall((:sigops_last_error, libsignalops), Cstring, (Ptr{Cvoid},), C_NULL)))
        s = new(devs, ctx[], Dict{Vector{UInt8},Ptr{Cvoid}}())
        finalizer(s) do s
            foreach(p -> ccall((:sigops_plan_destroy, libsignalops), Cvoid, (Ptr{Cvoid},), p), values(s.plans))
            ccall((:sigops_ctx_destroy, libsignalops), Cvoid, (Ptr{Cvoid},), s.ctx)
        end
    end
end

check(to::GPUSink, rc) = rc == 0 ||
    error(unsafe_string(ccall((:sigops_last_error, libsignalops), Cstring, (Ptr{Cvoid},), to.ctx)))

function compiled(to::GPUSink, bytes::Vector{UInt8})
    get!(to.plans, bytes) do
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(to, ccall((:sigops_plan_create, libsignalops), Cint,
                        (Ptr{Cvoid}, Ptr{UInt8}, Csize_t, Ref{Ptr{Cvoid}}), to.ctx, bytes, length(bytes), p))
        p[]
    end
end

dtypecode(::Type{Float32}) = Int32(1)
dtypecode(::Type{Float64}) = Int32(2)
dtypecode(::Type{<:Integer}) = Int32(3)
buffer(a::AbstractVecOrMat{T}) where T =
    SigopsBuffer(pointer(a), size(a, 1), size(a, 2), dtypecode(T), max(size(a, 1), stride(a, 2)))

# ---- the sink methods ---------------------------------------------------------------------
# Same shape as `sink(x,to::String)` (src/sink.jl:139-142): validate, hand to the backend,
# wrap like `initsink(x,T,data)` (src/sink.jl:120-121).
SignalOperators.sink(to::GPUSink) = x -> sink(x, to)
function SignalOperators.sink(x, to::GPUSink)
    x = process_sink_params(x)                                   # src/sink.jl:94-99
    plan = lower(x)                                              # graph -> stages (below)
    result = Array{sampletype(x),2}(undef, nframes(x), nchannels(x))   # initsink, src/sink.jl:115-117
    run!(to, plan, [result])
    initsink(x, refineroot(root(x)), result)
end

# `sink!(result, x)` semantics of src/sink.jl:158-168: a prefix of x, forced channel count.
function SignalOperators.sink!(result::Union{AbstractVector,AbstractMatrix}, x, to::GPUSink)
    nframes(x) < size(result, 1) && error("Signal is too short to fill buffer of length $(size(result,1)).")
    x = ToChannels(x, size(result, 2))
    run!(to, lower(x; nframes = size(result, 1), eltype = eltype(result)), [result])
    result
end

# Additive API: a batch of structurally identical graphs = one plan, many instances.
function SignalOperators.sink(xs::AbstractVector, to::GPUSink)
    xs = process_sink_params.(xs)
    plans = lower.(xs)
    all(p -> p.bytes == plans[1].bytes, plans) || error("batch elements do not lower to the same plan")
    results = [Array{sampletype(x),2}(undef, nframes(x), nchannels(x)) for x in xs]
    run!(to, plans, results)
    [initsink(x, refineroot(root(x)), r) for (x, r) in zip(xs, results)]
end

function run!(to::GPUSink, plans, results)
    plans = plans isa AbstractVector ? plans : [plans]
    handle = compiled(to, plans[1].bytes)
    ins = [buffer(a) for p in plans for a in p.inputs]
    outs = [buffer(r) for r in results]
    stats = SigopsStats()
    GC.@preserve plans results begin                              # caller owns every host buffer
        check(to, ccall((:sigops_plan_run, libsignalops), Cint,
                        (Ptr{Cvoid}, Int64, Ptr{SigopsBuffer}, Ptr{SigopsBuffer}, Ref{SigopsStats}),
                        handle, length(results), ins, outs, stats))
    end
    stats
end

# ---- lowering: the reference's node types -> plan stages ------------------------------------
# One method per node type of SURVEY.md Appendix E.  Each returns pieces
# (lo, hi, clo, chi, program) for consumer frames [lo,hi) / channels [clo,chi), where the
# node's own 0-based frame is n+shift and its channel is c*cm+co — the same recursion as
# Lowerer.lower in host/lowering.py, which is the executable specification of this code.
#
#   arrays / (array,fs)      LEAF_BUF  (src/arrays.jl:118-132)
#   NumberSignal             LEAF_CONST(x.val)                         (src/numbers.jl:62-64)
#   SignalFunction           LEAF_GEN  for `sin` and GPU-aware Functors (src/functions.jl:53-60);
#                            RandFn and arbitrary closures are evaluated by the CPU `sink`
#                            into an array first and become LEAF_BUF
#   CutApply Until           passes through; After shifts by resolvelen (src/cutting.jl:32,160-214)
#   PaddedSignal             split at nframes(child): child program | pad program (src/padding.jl:150-235)
#   AppendSignals            split at the cumulative child lengths          (src/appending.jl:92-110)
#   MapSignal                FnBr{+,*,-,/} fold left to right over `padded_signals`;
#                            ToEltypeFn -> CAST; AsNChannels/GetChanFn -> channel map;
#                            As1Channel -> sum of channel programs; tuplecat -> channel pieces
#                            (src/mapsignal.jl:219-272, src/reformatting.jl:148-184)
#   RampSignal               LEAF_RAMP_ON / LEAF_RAMP_OFF with L = resolvelen (src/ramps.jl:26,56-119)
#   NormedSignal             producer stage + sumsq slot; consumer: LEAF_BUF ./ LEAF_RMS (src/filters.jl:296-309)
#   FilteredSignal           h = x.fn(framerate(x)) on the host (DSP.jl, unchanged), then
#                              DF2TFilter-able  -> convert(SecondOrderSections, h): STAGE_IIR with
#                                                  [b0 b1 b2 a1 a2] per biquad and gain h.g
#                              FIRFilter        -> STAGE_FIR with pfb' / dpfb', inputDeficit, phiAccumulator
#                                                  or phiIdx read from h.kernel after setphase!
#                            (src/filters.jl:204-262, src/reformatting.jl:92-99)
#
# The byte layout (header, bufdescs, tabledescs, instrs, pieces, stages, Float64 blob) is
# documented in include/signalops.h; `write(io, htol(field))` per field in declaration order.

struct Plan
    bytes::Vector{UInt8}
    inputs::Vector{Array}
end

function lower(x; nframes = SignalOperators.nframes(x), eltype = sampletype(x))
    error("GPUSinks.lower: see host/lowering.py — port pending a Julia toolchain to test it against")
end

end # module
