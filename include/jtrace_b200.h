/* jtrace_b200.h -- C ABI of libjtrace_b200.so, the B200 drop-in for julia-raytracer's
 * per-pixel render loop.
 *
 * The reference has no FFI; the seam is the single call site
 *     src/jtrace.jl:85-94  ->  trace_samples(state, scene, bvh, lights, params, ...)   (src/trace.jl:215)
 * and the objects that call consumes. Every entry point below names the reference interface
 * it replaces. All functions return 0 on success or a negative jt_status; the message is
 * available from jt_last_error() (thread-local). No C++ exception, exit() or callback ever
 * crosses this boundary. The library never keeps a host pointer after a call returns.
 *
 * Conventions (identical to the Julia side, converted inside the library):
 *   - ids are 1-based, "none" is -1                      (src/scene.jl:45, :95-96)
 *   - vertex/element indices are Int64                   (src/math.jl:15-20)
 *   - the struct layouts are the ones Julia uses for the corresponding isbits structs, so a
 *     Vector{BvhNode}, Vector{InstanceData}, Vector{MaterialData}, ... is passed zero-copy
 *     with `pointer(v)` under GC.@preserve                
 *   - images are row-major, idx = W*j + i, RGBA linear float   (src/trace.jl:598, :631-648)
 */
#ifndef JTRACE_B200_H
#define JTRACE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JT_API __attribute__((visibility("default")))

typedef enum {
  JT_OK = 0,
  JT_ERR_INVALID = -1,     /* bad argument / inconsistent scene description */
  JT_ERR_CUDA = -2,        /* a CUDA runtime call failed */
  JT_ERR_NO_DEVICE = -3,   /* no CUDA device: there is no CPU fallback */
  JT_ERR_UNSUPPORTED = -4, /* feature that also crashes the reference */
  JT_ERR_PEER = -5,        /* multi-GPU group: a member device failed (message names the member) */
  JT_ERR_INTERNAL = -6
} jt_status;

/* ---- reference data layouts -------------------------------------------------------------- */

/* Frame3f = SVector{4,Vec3f}: columns x, y, z, o (src/math.jl:46) -- 48 B */
typedef struct { float x[3], y[3], z[3], o[3]; } jt_frame;

/* BvhNode (src/bvh.jl:34-39) -- 40 B */
typedef struct {
  float bbox_min[3], bbox_max[3];
  int64_t start;     /* 1-based: first child (children are start, start+1) or first primitive slot */
  int16_t num;
  int8_t axis;       /* 1..3 */
  uint8_t internal;  /* Bool */
  uint8_t _pad[4];
} jt_bvh_node;

/* BvhTree (src/bvh.jl:46-49) */
typedef struct {
  const jt_bvh_node* nodes;
  int64_t num_nodes;
  const int64_t* primitives; /* 1-based element / instance ids */
  int64_t num_primitives;
} jt_bvh_desc;

/* InstanceData (src/scene.jl:88-91) -- 64 B */
typedef struct { jt_frame frame; int64_t shape; int64_t material; } jt_instance;

/* MaterialData (src/scene.jl:213-229) -- 104 B; type = MaterialType enum 0..7 (src/scene.jl:191-200) */
typedef struct {
  int32_t type;
  float emission[3], color[3];
  float roughness, metallic, ior;
  float scattering[3];
  float scanisotropy, trdepth, opacity;
  int64_t emission_tex, color_tex, roughness_tex, scattering_tex, normal_tex;
} jt_material;

/* EnvironmentData (src/scene.jl:117-120) -- 72 B */
typedef struct { jt_frame frame; float emission[3]; int32_t _pad; int64_t emission_tex; } jt_environment;

/* CameraData without the name (src/scene.jl:48-56); not isbits in Julia -> filled field by field */
typedef struct {
  jt_frame frame;
  int32_t orthographic;
  float lens, film, aspect, focus, aperture;
} jt_camera;

/* TextureData (src/scene.jl:146-151): exactly one of pixelsf / pixelsb is non-NULL */
typedef struct {
  int64_t width, height;
  int32_t linear;
  int32_t _pad;
  const float* pixelsf;   /* Vec4f per texel, row-major, top row first */
  const uint8_t* pixelsb; /* Vec4b per texel */
} jt_texture_desc;

/* ShapeData (src/shape.jl:13-23) + its ShapeBvh (src/bvh.jl:51-55).
 * points / lines / radius / tangents are not accepted: they crash the reference. */
typedef struct {
  const float* positions;  int64_t num_positions;  /* Vec3f */
  const float* normals;    int64_t num_normals;    /* Vec3f, 0 = none */
  const float* texcoords;  int64_t num_texcoords;  /* Vec2f, 0 = none */
  const float* colors;     int64_t num_colors;     /* Vec4f, 0 = none */
  const int64_t* triangles; int64_t num_triangles; /* SVector{3,Int64}, 1-based */
  const int64_t* quads;     int64_t num_quads;     /* SVector{4,Int64}, 1-based */
  jt_bvh_desc bvh;
} jt_shape_desc;

/* TraceLight (src/trace.jl:102-105) */
typedef struct {
  int64_t instance;     /* 1-based or -1 */
  int64_t environment;  /* 1-based or -1 */
  const float* elements_cdf;
  int64_t num_elements;
} jt_light_desc;

/* SceneData + SceneBvh + TraceLights: everything trace_samples reads (src/trace.jl:215-224) */
typedef struct {
  int64_t num_cameras;      const jt_camera* cameras;
  int64_t num_instances;    const jt_instance* instances;
  int64_t num_environments; const jt_environment* environments;
  int64_t num_shapes;       const jt_shape_desc* shapes;
  int64_t num_textures;     const jt_texture_desc* textures;
  int64_t num_materials;    const jt_material* materials;
  int64_t num_lights;       const jt_light_desc* lights;
  jt_bvh_desc bvh;          /* SceneBvh.bvh: the TLAS over instances */
  /* srgb_to_rgb(b / 255f0) for b = 0..255 computed by the host (src/color.jl:12-23): only 256
   * distinct inputs exist, so a host-built table is bit-identical to the per-texel call in
   * lookup_texture (src/scene.jl:836-849). NULL = the library computes it in double. */
  const float* srgb_to_rgb_lut;
} jt_scene_desc;

/* Params fields read inside the hot path (src/cli.jl:90-108) + GPU extras */
typedef struct {
  int32_t camera;       /* 1-based index after find_camera (src/jtrace.jl:61) */
  int32_t resolution;
  int32_t samples;
  int32_t bounces;      /* 0..254 with the wavefront integrator (JT_ERR_INVALID otherwise) */
  int32_t sampler;      /* 1 = path, 2 = naive (src/cli.jl:88, :111-116) */
  int32_t clamp;        /* Params.clamp::Int (src/cli.jl:105) */
  int32_t nocaustics;
  int32_t envhidden;
  int32_t tentfilter;
  int32_t batch;
  int32_t bvhstacksize; /* accepted, unused: traversal stacks live in registers/local memory */
  int32_t traversal;    /* 0 = quantised wide BVH (fast), 1 = reference binary BVH, reference order */
  uint64_t seed;        /* counter-based RNG seed (jt_rng.h) */
  int32_t accumulate;   /* 0 = running-mean lerp exactly like src/trace.jl:631-647 (Q13)
                           1 = plain sums (divide at download) -- what multi-GPU sharding reduces */
  int32_t integrator;   /* 0 = wavefront (queues + compaction, default), 1 = one-thread-per-pixel megakernel */
  int32_t _reserved[6];
} jt_params;

/* Ray3f (src/geometry.jl:36-40) -- 32 B */
typedef struct { float o[3], d[3], tmin, tmax; } jt_ray;

/* SceneIntersection (src/shape.jl:61-66) -- 32 B; miss = {-1,-1,(0,0),0,false} */
typedef struct { int64_t instance, element; float uv[2]; float distance; uint8_t hit; uint8_t _pad[3]; } jt_hit;

/* Work counters accumulated by the render kernels since creation / last reset. */
typedef struct {
  uint64_t camera_paths;      /* trace_sample calls */
  uint64_t scene_rays;        /* intersect_scene_bvh calls   (src/trace.jl:298,490) */
  uint64_t light_rays;        /* intersect_instance_bvh calls (src/trace.jl:1025) */
  uint64_t kernel_launches;   /* CUDA kernels launched by the library */
  uint64_t extend_kernel_us;  /* device time of the dominant kernel (extend = closest hit), summed over its launches,
                                 CUDA events on the launching streams */
  uint64_t extend_launches;   /* number of extend launches timed */
  uint64_t stolen_samples;    /* samples traced by a path slot that started the range on another pixel (work stealing) */
  uint64_t resumed_rays;      /* closest-hit queries parked in a launch tail and resumed by the next launch */
} jt_counters;

typedef struct jt_scene jt_scene; /* device-resident scene: SceneData + SceneBvh + TraceLights */
typedef struct jt_state jt_state; /* device-resident TraceState (src/trace.jl:87-96) */

/* ---- library ------------------------------------------------------------------------------ */
JT_API const char* jt_last_error(void);
JT_API const char* jt_version(void);
JT_API int jt_device_count(void);

/* ---- scene: replaces handing (scene, bvh, lights) to trace_samples ------------------------- */
/* Copies the whole description to `device` and builds the wide BVH from the given binary BVH. */
JT_API int jt_scene_create(const jt_scene_desc* desc, int device, jt_scene** out);
JT_API void jt_scene_destroy(jt_scene* scene);
JT_API int jt_scene_counters(jt_scene* scene, jt_counters* out, int reset);
/* Structural statistics of the device layout (for DESIGN.md / bench reporting). */
typedef struct {
  int64_t wide_nodes, wide_node_bytes, prim_records, prim_record_bytes;
  int64_t inlined_instances, instanced_instances, texture_bytes, total_device_bytes;
  int64_t wide_depth_top;      /* depth of the top-level wide BVH */
  int64_t wide_depth_blas;     /* deepest instanced BLAS */
  int64_t opened_instances;    /* instances opened into the top-level tree (flattened / braided; INTEGRATION.md knobs) */
  int64_t wide_bvh_from_cache; /* 1 = the wide BVH came from the cache directory (jt_set_bvh_cache_dir), 0 = built */
  int64_t _reserved[4];
} jt_scene_stats;
JT_API int jt_scene_get_stats(jt_scene* scene, jt_scene_stats* out);

/* ---- state: make_trace_state (src/trace.jl:189-213) ---------------------------------------- */
JT_API int jt_state_create(jt_scene* scene, const jt_params* params, jt_state** out);
JT_API void jt_state_destroy(jt_state* state);
JT_API int jt_state_size(jt_state* state, int32_t* width, int32_t* height, int32_t* samples);
JT_API int jt_state_reset(jt_state* state);
/* Copy the accumulators to the host in the reference's layouts (state.image :: Vector{Vec4f},
 * albedo / normal :: Vector{Vec3f}, hits :: Vector{Int64}); any pointer may be NULL.
 * In accumulate = 1 mode the sums are divided by `state.samples` first. */
JT_API int jt_state_download(jt_state* state, float* image_rgba, float* albedo_rgb,
                             float* normal_rgb, int64_t* hits);
/* N3: the reference's save path on the GPU -- rgb_to_srgb (src/color.jl:25-29), clamp01nan and 8-bit
 * quantisation of save_image (src/sceneio.jl:97-113): width*height RGBA8, row-major, ready for the PNG writer. */
JT_API int jt_state_download_srgb8(jt_state* state, uint8_t* rgba8);
/* Raw device pointers of the accumulators (float4 image, float4-padded albedo / normal, int32
 * hits) so that a host that owns a communicator (torch.distributed / NCCL) can reduce them in
 * place across GPUs. `count` = width*height.
 * Ordering: the call flushes the lazily batched samples and BLOCKS until every kernel the library enqueued for this
 * state has completed, so the caller may reduce on any stream right away. The caller must in turn synchronise ITS
 * stream before the next library call on this state (jt_state_set_samples + jt_state_download read the buffers on
 * the library's own non-blocking stream). */
JT_API int jt_state_device_buffers(jt_state* state, void** image, void** albedo, void** normal,
                                   void** hits, int64_t* count);
/* After an external sum-reduction across ranks: set the number of samples the buffers hold. */
JT_API int jt_state_set_samples(jt_state* state, int32_t samples);

/* ---- the hot path --------------------------------------------------------------------------- */
/* trace_samples (src/trace.jl:215-274): advances state.samples by params.batch (clamped to
 * params.samples). Lazily batched: the reference calls this samples/batch times with batch = 1 by default,
 * so contiguous requests are merged and launched in chunks of 512 samples (JT_LAZY_SPP) or at the next
 * synchronisation point (download, jt_synchronize, counters, device_buffers). Nothing is copied back. */
JT_API int jt_trace_samples(jt_scene* scene, jt_state* state, const jt_params* params);
/* Same loop body for an explicit range of global sample indices [begin, end): the unit that
 * is sharded across GPUs. Does not touch state.samples bookkeeping beyond adding end-begin. */
JT_API int jt_trace_sample_range(jt_scene* scene, jt_state* state, const jt_params* params,
                                 int32_t sample_begin, int32_t sample_end);
/* Block until all enqueued work on this scene's stream is complete. */
JT_API int jt_synchronize(jt_scene* scene);
/* Device time in milliseconds spent by the render kernels enqueued since the last call
 * (CUDA events on the library's stream); synchronises. */
JT_API int jt_elapsed_ms(jt_scene* scene, float* ms);

/* ---- parity hooks ("identical rays") -------------------------------------------------------- */
/* intersect_scene_bvh (src/bvh.jl:306-371) for n host rays; find_any = false.
 * traversal: 0 = wide BVH kernel, 1 = reference-order kernel. Results use 1-based ids. */
JT_API int jt_intersect(jt_scene* scene, const jt_ray* rays, int64_t n, int traversal, jt_hit* out);
/* intersect_instance_bvh (src/bvh.jl:493-520): one named instance (1-based) per ray. */
JT_API int jt_intersect_instance(jt_scene* scene, const jt_ray* rays, const int64_t* instances,
                                 int64_t n, int traversal, jt_hit* out);
/* sample_camera + eval_camera (src/trace.jl:651-674, src/scene.jl:372-411) for explicit
 * (i, j, puv, luv) tuples: rays out. */
JT_API int jt_sample_camera(jt_scene* scene, const jt_params* params, int32_t width, int32_t height,
                            const int32_t* ij, const float* puv_luv, int64_t n, jt_ray* out);
/* Device-resident variant used by the benchmark: rays and hits are device pointers. */
JT_API int jt_intersect_device(jt_scene* scene, const void* d_rays, int64_t n, int traversal,
                               void* d_hits);

/* ---- multi-GPU group: ONE host thread drives N devices --------------------------------------
 * The reference's call site (src/jtrace.jl:83-94) is single-threaded, so sharding lives inside the library: the scene
 * is staged once and replicated on every member device, each member has a worker thread with its own streams, a
 * requested range of global sample indices is split into contiguous sub-ranges (one per member; RNG streams are keyed
 * by the global sample index, so the union is exactly the single-GPU sample set), members accumulate plain SUMS, and
 * the download runs ONE fused kernel on member 0 that reads every member's four sum buffers over NVLink peer access
 * (P2P loads; 52 B per pixel per remote member), adds them in member order, divides by the sample count and packs the
 * reference's host layouts. No NCCL is involved: 8 x 14.7 MB does not need a ring. `devices` may name a device more
 * than once (logical shards on one GPU; that is how the path is tested on a single-GPU box).
 * All group calls are asynchronous except create / synchronize / download / counters; calls on one group are not
 * re-entrant. A member failure surfaces as JT_ERR_PEER (or the member's own status) from the next blocking call. */
typedef struct jt_group jt_group;
typedef struct jt_group_state jt_group_state;
typedef struct {
  int32_t members;             /* number of member (device, scene) pairs */
  int32_t distinct_devices;
  int32_t peer_members;        /* members whose buffers member 0 reads through peer access (NVLink / PCIe P2P) */
  int32_t staged_members;      /* members without peer access: copied into a staging buffer on member 0's device first */
  int64_t reduce_bytes_remote; /* bytes read from other devices by the downloads so far */
  int64_t downloads;
  double stage_seconds;        /* host: flatten + wide-BVH build (once for the whole group) */
  double upload_seconds;       /* wall time of the N parallel uploads */
  int64_t _reserved[4];
} jt_group_stats;
JT_API int jt_group_create(const jt_scene_desc* desc, const int* devices, int n, jt_group** out);
JT_API void jt_group_destroy(jt_group* group);
JT_API int jt_group_get_stats(jt_group* group, jt_group_stats* out);
/* The member's scene, for the parity hooks and per-device counters (borrowed: do not destroy). */
JT_API int jt_group_scene(jt_group* group, int member, jt_scene** out);
/* Counters summed over the members (kernel time fields: maximum over members). Blocks until the group is idle. */
JT_API int jt_group_counters(jt_group* group, jt_counters* out, int reset);
/* make_trace_state for the group: one sum-mode state per member + the merged host-facing bookkeeping. */
JT_API int jt_group_state_create(jt_group* group, const jt_params* params, jt_group_state** out);
JT_API void jt_group_state_destroy(jt_group_state* state);
JT_API int jt_group_state_size(jt_group_state* state, int32_t* width, int32_t* height, int32_t* samples);
JT_API int jt_group_state_reset(jt_group_state* state);
/* trace_samples (src/trace.jl:215-274) across the group: same bookkeeping as jt_trace_samples; requests are merged
 * until members * 512 samples are pending or a blocking call arrives, then split over the members. Returns at once. */
JT_API int jt_group_trace_samples(jt_group* group, jt_group_state* state, const jt_params* params);
JT_API int jt_group_trace_sample_range(jt_group* group, jt_group_state* state, const jt_params* params,
                                       int32_t sample_begin, int32_t sample_end);
/* Block until every member is idle; returns the first member error, if any. */
JT_API int jt_group_synchronize(jt_group* group);
/* Merged download: same host layouts as jt_state_download / jt_state_download_srgb8 (fused P2P reduce + finalize). */
JT_API int jt_group_state_download(jt_group_state* state, float* image_rgba, float* albedo_rgb, float* normal_rgb,
                                   int64_t* hits);
JT_API int jt_group_state_download_srgb8(jt_group_state* state, uint8_t* rgba8);

/* ---- diagnostics ------------------------------------------------------------------------------------------------ */
/* Read bandwidth of a `bytes`-sized device buffer streamed `reps` times by one resident wave of 128-bit loads (best of
 * 5 launches, CUDA events), in GB/s. With bytes well below the L2 size this is the L2 roofline denominator SURVEY.md 8d
 * asks for; with bytes >> L2 it reproduces the HBM read peak. */
JT_API int jt_probe_read_bandwidth(int device, int64_t bytes, int reps, float* gbs_out);

/* ---- host-side helpers (CPU code; the steps bvh.jl performs on the Julia host) --------------- */
/* make_bvh (src/bvh.jl:138-183) with split_middle (:185-216) or split_sah (:218-274):
 * bboxes = n x {min[3], max[3]} floats. nodes_out must hold 2*n+1 entries, primitives_out n. */
JT_API int jt_make_bvh(const float* bboxes, int64_t n, int high_quality, jt_bvh_node* nodes_out,
                       int64_t* num_nodes_out, int64_t* primitives_out);

/* N4: make_trace_lights (src/trace.jl:117-187) on the GPU. Element weights (shape-local triangle / quad
 * areas, max(texel) * sin(theta) per environment texel) by one thread per element; the CDF stays the reference's
 * SEQUENTIAL Float32 prefix sum (a warp replays the host loop's additions in order), so the arrays are bit-identical to
 * the host builders'. `desc` needs cameras / lights / bvh not to be set. The returned descriptors (valid until
 * jt_lights_destroy) go into jt_scene_desc.lights. Flags: JT_LIGHTS_ENV_LUMINANCE weights environment texels by
 * max(R, G, B) instead of the reference's max(R, G, B, A = 1) (quirk Q8: with texels <= 1 the reference's CDF ignores
 * the image): unbiased, lower variance under a bright sun, NOT the reference's sample set. */
#define JT_LIGHTS_ENV_LUMINANCE 1
typedef struct jt_lights jt_lights;
JT_API int jt_lights_create(const jt_scene_desc* desc, int device, int flags, jt_lights** out);
JT_API int jt_lights_desc(jt_lights* lights, const jt_light_desc** descs, int64_t* count);
JT_API void jt_lights_destroy(jt_lights* lights);

/* N1: where jt_scene_create / jt_group_create keep finished wide BVHs (the expensive host step for
 * instancing-heavy scenes: 4.7 s for ecosys' 16.8 M flattened records). Files are named by a hash of everything the
 * builder reads (src/bvh.jl:66-304 replaced: shapes, instance frames, the host's BVHs, tuning knobs), written atomically,
 * and ignored when they do not match. NULL or "" turns caching off (the default, unless JT_BVH_CACHE_DIR is set). */
JT_API int jt_set_bvh_cache_dir(const char* dir);

/* The whole host side natively (SURVEY.md 8f N2 and the host halves of N1 / N4), for hosts that are not Julia:
 *   jt_host_scene_load   = load_scene (src/sceneio.jl:25-93): the scene JSON, PLY shapes (src/shape.jl:78-124, quad
 *                          promotion / fan triangulation :302-446, v-flip), 8-bit PNG and Radiance .hdr textures
 *                          (src/scene.jl:164-189, with the reference loader's HDR rule) and the missing-asset rule
 *   jt_host_scene_build  = make_scene_bvh (src/bvh.jl:66-136) + make_trace_lights (src/trace.jl:117-187)
 *   jt_host_scene_desc   = the flattening pass: a jt_scene_desc over arrays owned by the handle (valid until destroy;
 *                          BVHs and lights are present after jt_host_scene_build)
 * Bit-identical to the Python mirror (sceneio.py / bvh.py / lights.py) on every shipped scene: tests/test_native_host.py. */
typedef struct jt_host_scene jt_host_scene;
JT_API int jt_host_scene_load(const char* json_path, jt_host_scene** out);
JT_API int jt_host_scene_build(jt_host_scene* scene, int high_quality_bvh);
JT_API int jt_host_scene_desc(jt_host_scene* scene, const jt_scene_desc** out);
/* find_camera (src/scene.jl:358-370): 1-based index of `name`, else of "default" / "camera" / "camera0" / "camera1",
 * else 1; -1 if the scene has no camera. */
JT_API int jt_host_scene_find_camera(jt_host_scene* scene, const char* name, int32_t* camera);
/* Substitutions made by the missing-asset rule (one human-readable line each). */
JT_API int jt_host_scene_num_notes(jt_host_scene* scene);
JT_API const char* jt_host_scene_note(jt_host_scene* scene, int index);
JT_API void jt_host_scene_destroy(jt_host_scene* scene);

#ifdef __cplusplus
}
#endif
#endif /* JTRACE_B200_H */
