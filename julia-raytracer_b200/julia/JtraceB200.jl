#=
JtraceB200: ccall glue between julia-raytracer's host code and libjtrace_b200.so.

The Julia host keeps loading the scene (sceneio.jl), building the BVH (bvh.jl) and the lights
(trace.jl:117); this module adds the flattening pass and replaces ONE function of the hot path:

    trace_samples(state, scene, bvh, lights, params, bvh_stacks, bvh_sub_stacks, volume_stacks)
                                                                              (src/trace.jl:215-224)

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
end

check(rc) = rc == 0 || error("libjtrace_b200: ", unsafe_string(ccall((:jt_last_error, LIB), Cstring, ())))

# ---- the new flattening pass ------------------------------------------------------------------------
mutable struct GpuScene
    h::Ptr{Cvoid}
    keep::Vector{Any}    # everything the description pointed into (alive until the upload returned)
end

ptr_or_null(v::Vector{T}) where {T} = isempty(v) ? Ptr{T}(C_NULL) : pointer(v)

function gpu_scene(scene::SceneData, bvh::SceneBvh, lights::TraceLights; device::Integer = 0)::GpuScene
    check_layouts()
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
    keep = Any[scene, bvh, lights, cams, shapes, texs, lts, lut]
    GC.@preserve keep check(ccall((:jt_scene_create, LIB), Cint, (Ref{JtSceneDesc}, Cint, Ref{Ptr{Cvoid}}),
                                  desc, device, h))
    g = GpuScene(h[], Any[])
    finalizer(x -> ccall((:jt_scene_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h), g)
    g
end

# ---- TraceState with a device twin --------------------------------------------------------------------
mutable struct GpuState
    host::TraceState          # the reference's own struct: image / albedo / normal / hits / samples
    h::Ptr{Cvoid}
    scene::GpuScene
end

function gpu_state(g::GpuScene, make_trace_state::Function, scene::SceneData, params::Params)::GpuState
    host = make_trace_state(scene, params)            # src/trace.jl:189-213 (sizes + zeroed buffers)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jt_state_create, LIB), Cint, (Ptr{Cvoid}, Ref{JtParams}, Ref{Ptr{Cvoid}}), g.h, JtParams(params), h))
    st = GpuState(host, h[], g)
    finalizer(x -> ccall((:jt_state_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h), st)
    st
end

function sync!(st::GpuState)
    s = st.host
    GC.@preserve s check(ccall((:jt_state_download, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Vec4f}, Ptr{Vec3f}, Ptr{Vec3f}, Ptr{Int64}),
        st.h, pointer(s.image), pointer(s.albedo), pointer(s.normal), pointer(s.hits)))
    st
end

"""
    trace_samples(state, scene, bvh, lights, params, bvh_stacks, bvh_sub_stacks, volume_stacks)

Drop-in for `Trace.trace_samples` (src/trace.jl:215-274): `state` is a `GpuState`; `scene`, `bvh`,
`lights` already live on the device inside `state.scene`; the scratch stacks are ignored.
"""
function trace_samples(st::GpuState, scene, bvh, lights, params::Params, bvh_stacks = nothing,
                       bvh_sub_stacks = nothing, volume_stacks = nothing)
    st.host.samples >= params.samples && return
    check(ccall((:jt_trace_samples, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{JtParams}), st.scene.h, st.h, JtParams(params)))
    st.host.samples = min(st.host.samples + params.batch, params.samples)
    st.host.samples == params.samples && sync!(st)       # get_image(state.host) then works unchanged
    nothing
end

end # module
