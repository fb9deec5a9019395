#=
JtraceB200: ccall glue between julia-raytracer's host code and libjtrace_b200.so.

The Julia host keeps loading the scene (sceneio.jl), building the BVH (bvh.jl), the lights
(trace.jl:117) and the TraceState (trace.jl:189); this module adds the flattening pass and ONE new
method of the hot-path function, selected by dispatch on the type of its `scene` argument:

    trace_samples(state::TraceState, scene::GpuScene, bvh, lights, params, bvh_stacks, bvh_sub_stacks, volume_stacks)
                                                                              (src/trace.jl:215-224)

`state` stays the reference's own TraceState: `state.samples` advances exactly like the reference's
(src/trace.jl:225-232, read by the progress print at src/jtrace.jl:96-105), and its image / albedo /
normal / hits vectors are refreshed from the device on the final call, so `get_image(state)`
(src/trace.jl:676) and `save_image` run unchanged. INTEGRATION.md has the diff of src/jtrace.jl.

NOTE: written without a Julia toolchain at hand (none in the build image, no network). Struct layouts
follow SURVEY.md Appendix B; `check_layouts()` asserts them against `sizeof`/`fieldoffset` on first use.
The Python mirror (julia-raytracer_b200/flatten.py, trace.py) passes the very same flat buffers through
the same C ABI and is what the test-suite exercises.
=#
module JtraceB200

using ..Math: Vec2f, Vec3f, Vec4f, Vec4b, Frame3f
using ..Scene: SceneData, CameraData, InstanceData, MaterialData, EnvironmentData, TextureData
using ..Shape: ShapeData
using ..Bvh: SceneBvh, BvhNode, BvhTree
using ..Trace: TraceLights, TraceLight, TraceState
import ..Trace: trace_samples          # extended with a method for GpuScene below
using ..Color: srgb_to_rgb
using ..Cli: Params

const LIB = joinpath(@__DIR__, "..", "libjtrace_b200.so")

# ---- mirrors of include/jtrace_b200.h -------------------------------------------------------------
struct JtBvhDesc
    nodes::Ptr{BvhNode}
    num_nodes::Int64
    primitives::Ptr{Int64}
    num_primitives::Int64
end
JtBvhDesc(t::BvhTree) = JtBvhDesc(pointer(t.nodes), length(t.nodes), pointer(t.primitives), length(t.primitives))

struct JtShapeDesc
    positions::Ptr{Vec3f};  num_positions::Int64
    normals::Ptr{Vec3f};    num_normals::Int64
    texcoords::Ptr{Vec2f};  num_texcoords::Int64
    colors::Ptr{Vec4f};     num_colors::Int64
    triangles::Ptr{Int64};  num_triangles::Int64
    quads::Ptr{Int64};      num_quads::Int64
    bvh::JtBvhDesc
end

struct JtTextureDesc
    width::Int64
    height::Int64
    linear::Int32
    _pad::Int32
    pixelsf::Ptr{Vec4f}
    pixelsb::Ptr{Vec4b}
end

struct JtLightDesc
    instance::Int64
    environment::Int64
    elements_cdf::Ptr{Float32}
    num_elements::Int64
end

struct JtCamera
    frame::Frame3f
    orthographic::Int32
    lens::Float32
    film::Float32
    aspect::Float32
    focus::Float32
    aperture::Float32
end
JtCamera(c::CameraData) = JtCamera(c.frame, c.orthographic ? 1 : 0, c.lens, c.film, c.aspect, c.focus, c.aperture)

struct JtSceneDesc
    num_cameras::Int64;      cameras::Ptr{JtCamera}
    num_instances::Int64;    instances::Ptr{InstanceData}
    num_environments::Int64; environments::Ptr{EnvironmentData}
    num_shapes::Int64;       shapes::Ptr{JtShapeDesc}
    num_textures::Int64;     textures::Ptr{JtTextureDesc}
    num_materials::Int64;    materials::Ptr{MaterialData}
    num_lights::Int64;       lights::Ptr{JtLightDesc}
    bvh::JtBvhDesc
    srgb_to_rgb_lut::Ptr{Float32}
end

struct JtParams
    camera::Int32; resolution::Int32; samples::Int32; bounces::Int32; sampler::Int32; clamp::Int32
    nocaustics::Int32; envhidden::Int32; tentfilter::Int32; batch::Int32; bvhstacksize::Int32
    traversal::Int32
    seed::UInt64
    accumulate::Int32
    integrator::Int32
    _r1::Int32; _r2::Int32; _r3::Int32; _r4::Int32; _r5::Int32; _r6::Int32
end
JtParams(p::Params; traversal = 0, seed = 0, accumulate = 0, integrator = 0) = JtParams(
    p.camera, p.resolution, p.samples, p.bounces, p.sampler, p.clamp, p.nocaustics, p.envhidden,
    p.tentfilter, p.batch, p.bvhstacksize, traversal, seed, accumulate, integrator, 0, 0, 0, 0, 0, 0)

function check_layouts()
    @assert sizeof(BvhNode) == 40 && fieldoffset(BvhNode, 2) == 24 && fieldoffset(BvhNode, 3) == 32
    @assert fieldoffset(BvhNode, 4) == 34 && fieldoffset(BvhNode, 5) == 35
    @assert sizeof(InstanceData) == 64 && fieldoffset(InstanceData, 2) == 48
    @assert sizeof(MaterialData) == 104 && fieldoffset(MaterialData, 11) == 64
    @assert sizeof(EnvironmentData) == 72 && fieldoffset(EnvironmentData, 3) == 64
    @assert sizeof(Frame3f) == 48 && sizeof(Vec4f) == 16 && sizeof(Vec4b) == 4
    @assert sizeof(JtParams) == 88 && sizeof(JtBvhDesc) == 32 && sizeof(JtShapeDesc) == 128
    @assert sizeof(JtTextureDesc) == 40 && sizeof(JtLightDesc) == 32 && sizeof(JtCamera) == 72
    @assert sizeof(JtSceneDesc) == 152
end

check(rc) = rc == 0 || error("libjtrace_b200: ", unsafe_string(ccall((:jt_last_error, LIB), Cstring, ())))

# ---- the new flattening pass + the device twin of the TraceState --------------------------------------
mutable struct GpuScene
    h::Ptr{Cvoid}          # jt_scene* (one device) or jt_group* (several devices, sharded inside the library)
    state::Ptr{Cvoid}      # jt_state* / jt_group_state*: device twin of the host TraceState
    grouped::Bool
    host::TraceState       # the reference's own state: refreshed by sync!
end

ptr_or_null(v::Vector{T}) where {T} = isempty(v) ? Ptr{T}(C_NULL) : pointer(v)

function destroy!(g::GpuScene)
    if g.grouped
        g.state != C_NULL && ccall((:jt_group_state_destroy, LIB), Cvoid, (Ptr{Cvoid},), g.state)
        g.h != C_NULL && ccall((:jt_group_destroy, LIB), Cvoid, (Ptr{Cvoid},), g.h)
    else
        g.state != C_NULL && ccall((:jt_state_destroy, LIB), Cvoid, (Ptr{Cvoid},), g.state)
        g.h != C_NULL && ccall((:jt_scene_destroy, LIB), Cvoid, (Ptr{Cvoid},), g.h)
    end
    g.state = C_NULL
    g.h = C_NULL
    nothing
end

"""
    gpu_scene(scene, bvh, lights, state, params; devices = [0], traversal = 0, seed = 0)

Flatten (scene, bvh, lights) into the C description, upload it (replicated on every device of `devices`)
and create the device twin of `state` (the TraceState `make_trace_state` returned). Called once, after
`make_trace_state` (src/jtrace.jl:69). The library copies everything before returning.
"""
function gpu_scene(scene::SceneData, bvh::SceneBvh, lights::TraceLights, state::TraceState, params::Params;
                   devices::Vector{<:Integer} = [0], traversal::Integer = 0, seed::Integer = 0,
                   bvh_cache_dir::AbstractString = "")::GpuScene
    check_layouts()
    # N1: keep finished wide BVHs on disk (hash-named, atomically written); "" leaves caching to JT_BVH_CACHE_DIR
    bvh_cache_dir == "" || check(ccall((:jt_set_bvh_cache_dir, LIB), Cint, (Cstring,), bvh_cache_dir))
    cams = [JtCamera(c) for c in scene.cameras]
    # vertex index vectors are Vector{SVector{k,Int64}}: reinterpret as flat Int64
    shapes = JtShapeDesc[]
    for (s, sh) in enumerate(scene.shapes)
        push!(shapes, JtShapeDesc(
            ptr_or_null(sh.positions), length(sh.positions), ptr_or_null(sh.normals), length(sh.normals),
            ptr_or_null(sh.texcoords), length(sh.texcoords), ptr_or_null(sh.colors), length(sh.colors),
            Ptr{Int64}(ptr_or_null(sh.triangles)), length(sh.triangles),
            Ptr{Int64}(ptr_or_null(sh.quads)), length(sh.quads), JtBvhDesc(bvh.shapes[s].bvh)))
    end
    texs = [JtTextureDesc(t.width, t.height, t.linear ? 1 : 0, 0, ptr_or_null(t.pixelsf), ptr_or_null(t.pixelsb))
            for t in scene.textures]
    lts = [JtLightDesc(l.instance, l.environment, ptr_or_null(l.elements_cdf), length(l.elements_cdf))
           for l in lights.lights]
    lut = Float32[srgb_to_rgb(b / 255.0f0) for b in 0:255]   # src/color.jl:18-23, 256 distinct inputs
    desc = JtSceneDesc(
        length(cams), pointer(cams), length(scene.instances), ptr_or_null(scene.instances),
        length(scene.environments), ptr_or_null(scene.environments), length(shapes), ptr_or_null(shapes),
        length(texs), ptr_or_null(texs), length(scene.materials), ptr_or_null(scene.materials),
        length(lts), ptr_or_null(lts), JtBvhDesc(bvh.bvh), pointer(lut))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    st = Ref{Ptr{Cvoid}}(C_NULL)
    jp = JtParams(params; traversal = traversal, seed = seed)
    keep = Any[scene, bvh, lights, cams, shapes, texs, lts, lut]
    grouped = length(devices) > 1
    if grouped
        devs = Cint[d for d in devices]
        GC.@preserve keep devs check(ccall((:jt_group_create, LIB), Cint,
            (Ref{JtSceneDesc}, Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}), desc, devs, length(devs), h))
        check(ccall((:jt_group_state_create, LIB), Cint, (Ptr{Cvoid}, Ref{JtParams}, Ref{Ptr{Cvoid}}), h[], jp, st))
    else
        GC.@preserve keep check(ccall((:jt_scene_create, LIB), Cint, (Ref{JtSceneDesc}, Cint, Ref{Ptr{Cvoid}}),
                                      desc, devices[1], h))
        check(ccall((:jt_state_create, LIB), Cint, (Ptr{Cvoid}, Ref{JtParams}, Ref{Ptr{Cvoid}}), h[], jp, st))
    end
    g = GpuScene(h[], st[], grouped, state)
    w = Ref{Int32}(0); ht = Ref{Int32}(0)
    if grouped
        check(ccall((:jt_group_state_size, LIB), Cint, (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}, Ptr{Int32}), g.state, w, ht, C_NULL))
    else
        check(ccall((:jt_state_size, LIB), Cint, (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}, Ptr{Int32}), g.state, w, ht, C_NULL))
    end
    @assert (w[], ht[]) == (state.width, state.height)     # make_trace_state's sizes (src/trace.jl:189-197)
    finalizer(destroy!, g)
    g
end

"Refresh the host TraceState (image / albedo / normal / hits, the reference's layouts) from the device."
function sync!(g::GpuScene)
    s = g.host
    GC.@preserve s begin
        if g.grouped
            check(ccall((:jt_group_state_download, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Vec4f}, Ptr{Vec3f}, Ptr{Vec3f}, Ptr{Int64}),
                g.state, pointer(s.image), pointer(s.albedo), pointer(s.normal), pointer(s.hits)))
        else
            check(ccall((:jt_state_download, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Vec4f}, Ptr{Vec3f}, Ptr{Vec3f}, Ptr{Int64}),
                g.state, pointer(s.image), pointer(s.albedo), pointer(s.normal), pointer(s.hits)))
        end
    end
    g
end

"""
    trace_samples(state, scene::GpuScene, bvh, lights, params, bvh_stacks, bvh_sub_stacks, volume_stacks)

The GPU method of `Trace.trace_samples` (src/trace.jl:215-274): same arguments in the same order; `bvh`,
`lights` and the three scratch stacks are ignored (they already live on the device / are not needed).
Enqueues `params.batch` more samples per pixel and returns; on the call that reaches `params.samples`
the host arrays of `state` are refreshed.
"""
function trace_samples(state::TraceState, g::GpuScene, bvh, lights, params::Params, bvh_stacks = nothing,
                       bvh_sub_stacks = nothing, volume_stacks = nothing)
    state === g.host || error("JtraceB200.trace_samples: `state` is not the TraceState given to gpu_scene")
    state.samples >= params.samples && return            # src/trace.jl:225-227
    jp = JtParams(params)
    if g.grouped
        check(ccall((:jt_group_trace_samples, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{JtParams}), g.h, g.state, jp))
    else
        check(ccall((:jt_trace_samples, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{JtParams}), g.h, g.state, jp))
    end
    state.samples = min(state.samples + max(params.batch, 1), params.samples)   # src/trace.jl:228-232
    state.samples == params.samples && sync!(g)
    nothing
end

end # module
