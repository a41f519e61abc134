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

# ---- plan bytes (include/signalops.h; field order as in host/lowering.py `Plan.tobytes`) ----------------
const MAGIC = 0x504F4753; const PLAN_VERSION = UInt32(1)
const OP_LOAD, OP_MUL, OP_CAST_F32 = UInt8(1), UInt8(4), UInt8(12)
const LEAF_NONE, LEAF_CONST, LEAF_BUF, LEAF_STAGE = UInt8(0), UInt8(1), UInt8(2), UInt8(8)
const STAGE_IIR, STAGE_FIR = Int32(2), Int32(3)
const FIR_ARBITRARY, FIR_RATIONAL, FIR_DECIMATOR = Int32(1), Int32(2), Int32(3)

struct Instr                      # sigops_instr, 80 bytes
    op::UInt8; leaf::UInt8; fn::UInt8; flags::UInt8
    buf::Int32; c_mul::Int32; c_off::Int32
    i0::Int64; i1::Int64; i2::Int64
    d0::Float64; d1::Float64; d2::Float64; d3::Float64; d4::Float64
end
Instr(op, leaf; buf = 0, i1 = 0, d0 = 0.0) =
    Instr(op, leaf, 0x00, 0x00, Int32(buf), Int32(1), Int32(0), 0, Int64(i1), 0, Float64(d0), 0.0, 0.0, 0.0, 0.0)
put(io, x::Instr) = foreach(f -> write(io, htol(getfield(x, f))), fieldnames(Instr))

Base.@kwdef struct Stage          # sigops_stage, 128 bytes
    kind::Int32; out_buf::Int32; sumsq_slot::Int32 = -1
    piece_start::Int32 = 0; n_pieces::Int32 = 0
    in_prog_start::Int32; in_prog_len::Int32; epi_prog_start::Int32; epi_prog_len::Int32
    nchannels::Int32; n_in::Int64; n_out::Int64
    n_sections::Int32 = 0; coef_table::Int32 = -1; gain::Float64 = 1.0
    fir_kind::Int32 = 0; n_phases::Int32 = 0; taps_per_phase::Int32 = 0
    pfb_table::Int32 = -1; dpfb_table::Int32 = -1; interpolation::Int32 = 0; decimation::Int32 = 0
    reserved0::Int32 = 0; input_deficit::Int64 = 0; rate::Float64 = 0.0; phase0::Float64 = 0.0
end
put(io, x::Stage) = foreach(f -> write(io, htol(getfield(x, f))), fieldnames(Stage))

function planbytes(bufs, tables::Vector{Vector{Float64}}, instrs::Vector{Instr}, stages::Vector{Stage},
                   n_inputs, n_temps, n_outputs)
    io = IOBuffer()
    blob = reduce(vcat, tables; init = Float64[])
    foreach(v -> write(io, htol(UInt32(v))),
            (MAGIC, PLAN_VERSION, n_inputs, n_temps, n_outputs, 0, length(tables), length(instrs), 0, length(stages)))
    write(io, htol(UInt64(length(blob))))
    for (n, c, dt) in bufs                       # sigops_bufdesc: inputs, temps, outputs
        write(io, htol(Int64(n))); write(io, htol(Int32(c))); write(io, htol(Int32(dt)))
    end
    off = 0
    for tb in tables                             # sigops_tabledesc
        write(io, htol(Int64(off))); write(io, htol(Int64(length(tb)))); off += length(tb)
    end
    foreach(i -> put(io, i), instrs)
    foreach(s -> put(io, s), stages)             # (no MAP stages here, hence no pieces)
    foreach(v -> write(io, htol(v)), blob)
    take!(io)
end

# ---- lowering ---------------------------------------------------------------------------------------------
# Implemented here for the two barrier shapes of the benchmark configurations — the ones whose whole
# cost is a kernel of this library:
#     (array, fs) |> Filt(...) [|> Amplify(number)]...        one STAGE_IIR with a gain epilogue
#     ToFramerate((array, fs), fs2)                            one STAGE_FIR
# Every other graph needs the general recursion of host/lowering.py (`Lowerer.lower`, ~700 lines of
# Python that this function should be a transcription of); it is reported as such rather than guessed at.
# UNTESTED: written against src/filters.jl:96-107, src/mapsignal.jl:8-17,131-145,183-186,
# src/numbers.jl:1-4, src/reformatting.jl:92-122 and SURVEY.md App. B without a Julia toolchain.
isarraysignal(x) = x isa Tuple{<:AbstractArray,<:Number}

function lower(x; nframes = SignalOperators.nframes(x), eltype = sampletype(x))
    # peel `Amplify(number)` layers: MapSignal(FnBr(*), ...) over (signal, NumberSignal...)  (src/mapsignal.jl:131-145)
    gains = Float64[]
    r = x
    while r isa MapSignal && r.fn isa FnBr && r.fn.fn === (*) && r.bychannel &&
          all(s -> s isa NumberSignal, Base.tail(r.signals))
        prepend!(gains, Float64[s.val for s in Base.tail(r.signals)])   # NumberSignal.val is already 10^(dB/20)
        r = first(r.signals)
    end
    (r isa FilteredSignal && isarraysignal(r.signal) && length(gains) <= 2) ||
        error("GPUSinks.lower: only `array |> Filt |> Amplify(number)` and `ToFramerate(array)` are transcribed ",
              "to Julia so far; see host/lowering.py for the general lowering of ", typeof(x))
    data = r.signal[1] isa AbstractVector ? reshape(r.signal[1], :, 1) : r.signal[1]
    T = Base.eltype(data)
    (T === Float64 || T === Float32) || error("GPUSinks.lower: sample type $T")
    nin, C = size(data)
    dt = dtypecode(T)
    h = r.fn(framerate(r))                                   # design at sink time, src/filters.jl:205
    load = Instr(OP_LOAD, LEAF_BUF; buf = 0, i1 = nin)       # zero padded past the input, src/filters.jl:240
    epi = Instr[Instr(OP_LOAD, LEAF_STAGE)]
    foreach(g -> push!(epi, Instr(OP_MUL, LEAF_CONST; d0 = g)), gains)
    T === Float32 && !isempty(gains) && push!(epi, Instr(OP_CAST_F32, LEAF_NONE))
    length(epi) == 1 && empty!(epi)
    if h isa DSP.Filters.FIRFilter                           # resampler, src/reformatting.jl:92-99
        isempty(gains) || error("GPUSinks.lower: gains after ToFramerate need the general lowering (separate MAP stage)")
        T === Float64 || error("GPUSinks.lower: Float32 resampling needs the widen/round stages of host/lowering.py")
        k = h.kernel
        common = (kind = STAGE_FIR, out_buf = Int32(1), in_prog_start = Int32(0), in_prog_len = Int32(1),
                  epi_prog_start = Int32(1), epi_prog_len = Int32(0), nchannels = Int32(C), n_in = nin, n_out = nframes,
                  input_deficit = Int64(k.inputDeficit))
        if k isa DSP.Filters.FIRArbitrary
            tables = [vec(collect(Float64, k.pfb)), vec(collect(Float64, k.dpfb))]      # column phi = [phase][tap]
            st = Stage(; common..., fir_kind = FIR_ARBITRARY, n_phases = Int32(k.Nϕ), taps_per_phase = Int32(k.tapsPerϕ),
                       pfb_table = Int32(0), dpfb_table = Int32(1), rate = Float64(k.rate), phase0 = Float64(k.ϕAccumulator))
        elseif k isa DSP.Filters.FIRRational || k isa DSP.Filters.FIRInterpolator
            tables = [vec(collect(Float64, k.pfb))]
            q = k isa DSP.Filters.FIRRational ? denominator(k.ratio) : 1
            st = Stage(; common..., fir_kind = FIR_RATIONAL, n_phases = Int32(k.Nϕ), taps_per_phase = Int32(k.tapsPerϕ),
                       pfb_table = Int32(0), interpolation = Int32(k.Nϕ), decimation = Int32(q), phase0 = Float64(k.ϕIdx))
        elseif k isa DSP.Filters.FIRDecimator
            tables = [collect(Float64, k.h)]                 # stored reversed by DSP.jl: window order
            st = Stage(; common..., fir_kind = FIR_DECIMATOR, n_phases = Int32(1), taps_per_phase = Int32(k.hLen),
                       pfb_table = Int32(0), interpolation = Int32(1), decimation = Int32(k.decimation), phase0 = 1.0)
        else
            error("GPUSinks.lower: single-rate FIR kernels are not lowered")
        end
        instrs = Instr[load]
    else                                                     # IIR: DF2T second-order sections, SURVEY.md App. B.2
        sos = convert(DSP.SecondOrderSections, h)
        M = length(sos.biquads)
        M <= 8 || error("GPUSinks.lower: cascades of more than 8 sections are split by host/lowering.py")
        coef = Float64[]
        for b in sos.biquads
            append!(coef, (b.b0, b.b1, b.b2, b.a1, b.a2))
        end
        tables = [coef]
        instrs = vcat(Instr[load], epi)
        st = Stage(kind = STAGE_IIR, out_buf = Int32(1), in_prog_start = Int32(0), in_prog_len = Int32(1),
                   epi_prog_start = Int32(1), epi_prog_len = Int32(length(epi)), nchannels = Int32(C), n_in = nframes,
                   n_out = nframes, n_sections = Int32(M), coef_table = Int32(0), gain = Float64(sos.g))
    end
    bufs = [(nin, C, dt), (nframes, C, dtypecode(eltype))]
    Plan(planbytes(bufs, tables, instrs, [st], 1, 0, 1), Array[data])
end

end # module
