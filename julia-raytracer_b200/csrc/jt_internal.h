// jt_internal.h -- internal declarations shared by the translation units of libjtrace_b200.so.
#pragma once
#include <stdint.h>

#include <vector_types.h>

#include <string>
#include <memory>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/jtrace_b200.h"

// ---- error channel (thread-local message, int status; nothing throws across the ABI) ---------
int jt_set_error(int code, const char* fmt, ...);

// ---- device scene layout -----------------------------------------------------------------------
// Everything the render loop reads lives in a handful of flat device arrays (all ids 0-based,
// -1 = none). Records that the traversal fetches are multiples of 16 B and 16 B-aligned so they
// are read with 128-bit loads.

struct JtShapeRec {      // 48 B
  int32_t kind;          // 0 = empty, 1 = triangles, 2 = quads
  int32_t num_elements;
  int32_t ref_node_off;  // first node of this shape's reference binary BVH in ref_nodes
  int32_t ref_prim_off;  // first slot in ref_prims
  int32_t elem_off;      // first element in elements (int4 each)
  int32_t pos_off;       // first vertex in positions (float3 each)
  int32_t norm_off;      // first vertex in normals or -1
  int32_t uv_off;        // first vertex in texcoords or -1
  int32_t col_off;       // first vertex in colors or -1
  int32_t wide_root;     // root node of the shape's wide BLAS in wnodes (-1 if the shape is empty)
  int32_t rank_off;      // first entry of this BLAS in each of the 8 octant rank tables
  int32_t num_ref_nodes;
};

struct JtInstanceRec {  // 112 B = 7 x float4
  float frame[12];      // x, y, z, o columns (src/math.jl:46)
  float inv[12];        // inverse(frame, true) (src/math.jl:95-110), precomputed (Q5: bit-identical)
  int32_t shape;
  int32_t material;
  int32_t inlined;      // 1 = identity frame, geometry lives in the top-level wide BVH
  int32_t _pad;
};

struct JtMaterialRec {  // 96 B
  int32_t type;
  float emission[3];
  float color[3];
  float roughness, metallic, ior;
  float scattering[3];
  float scanisotropy, trdepth, opacity;
  int32_t emission_tex, color_tex, roughness_tex, scattering_tex, normal_tex;
  int32_t _pad[3];
};

struct JtTextureRec {  // 32 B
  int32_t width, height;
  int32_t linear;
  int32_t is_float;
  int64_t offset;  // first texel in texels_f (float4) or texels_b (uchar4)
  int64_t _pad;
};

struct JtEnvRec {  // 64 B
  float frame[12];
  float emission[3];
  int32_t emission_tex;
};

struct JtLightRec {  // 32 B
  int32_t instance, environment;
  int32_t cdf_off, cdf_len;
  // guide table over the CDF's VALUE range (jt_stage.cpp: build_cdf_guide): bucket b = (int)(limit * guide_scale)
  // brackets upper_bound's answer between guide[b] + 1 and guide[b + 1] + 1. guide_len == 0: no table.
  int32_t guide_off, guide_len;
  float guide_scale;
  int32_t _pad;
};

struct JtCameraRec {  // 80 B
  float frame[12];
  int32_t orthographic;
  float lens, film, aspect, focus, aperture;
  int32_t _pad[2];
};

// Passed by value to every kernel.
struct JtDevScene {
  // reference (parity) traversal: the host-built binary BVHs, 32 B per node
  //   n[0] = {min.x, min.y, min.z, max.x}  n[1] = {max.y, max.z, bits(start), bits(num | axis<<16 | internal<<24)}
  const float4* ref_nodes;
  const int32_t* ref_prims;
  int32_t tlas_num_nodes;  // TLAS occupies ref_nodes[0 .. tlas_num_nodes)
  int32_t num_instances;
  // geometry
  const JtShapeRec* shapes;
  const float* positions;  // packed float3
  const float* normals;    // packed float3
  const float* texcoords;  // packed float2
  const float4* colors;
  const int4* elements;    // local 0-based vertex ids; triangles use xyz
  const JtInstanceRec* instances;
  // shading
  const JtMaterialRec* materials;
  const JtTextureRec* textures;
  const float4* texels_f;
  const uchar4* texels_b;
  const float* srgb_lut;  // 256 entries
  const JtEnvRec* environments;
  int32_t num_environments;
  int32_t num_lights;
  const JtLightRec* lights;
  const float* light_cdf;
  const int32_t* light_guide;  // guide_len + 1 entries per light that has a table
  const JtCameraRec* cameras;
  // wide (fast) traversal
  const float4* wnodes;    // 80 B (5 x float4) per node
  const float4* wtris;     // 48 B (3 x float4) per triangle record: {p1, elem} {e1, inst} {e2, flags}
  const uint32_t* tri_rank;   // [8][num_wtris]: reference visit rank per ray octant (tie-breaks)
  const uint32_t* inst_rank;  // [8][num_instances]
  const float4* inst_bounds;  // 2 per instance: padded world-space box {lo.xyz,-} {hi.xyz,-} (probe early-out)
  int32_t num_wtris;
  int32_t wide_root;       // root of the top-level wide BVH
};

// ---- host-side wide BVH builder (jt_wide_bvh.cpp) ------------------------------------------------
struct JtWideNode {  // 80 B, layout documented in jt_wide_bvh.cpp / DESIGN.md
  float p[3];
  uint8_t e[3];
  uint8_t imask;
  uint32_t child_base;
  uint32_t prim_base;
  uint8_t meta[8];
  uint8_t qlo[3][8];
  uint8_t qhi[3][8];
};
static_assert(sizeof(JtWideNode) == 80, "wide node must be 80 bytes");

struct JtWideTri {  // 48 B
  float p1[3];
  int32_t element;   // 0-based element id inside its shape
  float e1[3];
  int32_t instance;  // 0-based instance id for inlined geometry, -1 inside an instanced BLAS
  float e2[3];
  uint32_t flags;    // bit0: second half of a quad (uv -> 1-uv); bit8: instance leaf (enter the BLAS);
                     // bit9: triangle of a flattened instance (test in instance space: transform the ray first)
};
static_assert(sizeof(JtWideTri) == 48, "wide tri must be 48 bytes");

// ---- host-side staging of the scene description (jt_scene.cu) -> wide BVH builder ---------------
struct JtHostShape {
  int kind = 0;                 // 0 empty, 1 triangles, 2 quads
  std::vector<float> pos;       // 3 per vertex
  std::vector<int32_t> elems;   // 4 per element, 0-based (triangles: 4th = 3rd)
  std::vector<jt_bvh_node> ref_nodes;  // as given by the host (1-based fields)
  std::vector<int64_t> ref_prims;      // 1-based
  int64_t num_elements() const { return (int64_t)elems.size() / 4; }
};
struct JtHostInstance {
  float frame[12];
  float inv[12];
  int shape, material;
  bool inlined;
};
// std::vector whose resize() leaves trivially constructible elements uninitialised: the multi-GB arrays of a flattened
// scene are filled right after being sized (builder splice, cache read), and zeroing them first costs as much as the fill.
template <class T>
struct JtNoInitAlloc : std::allocator<T> {
  template <class U>
  struct rebind { using other = JtNoInitAlloc<U>; };
  template <class U>
  void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new ((void*)p) U; }
  template <class U, class... Args>
  void construct(U* p, Args&&... args) { ::new ((void*)p) U(std::forward<Args>(args)...); }
};
template <class T>
using JtBigVec = std::vector<T, JtNoInitAlloc<T>>;

struct JtWideResult {
  JtBigVec<JtWideNode> nodes;
  JtBigVec<JtWideTri> tris;
  JtBigVec<uint32_t> tri_rank[8];      // per octant, indexed like tris
  std::vector<uint32_t> inst_rank[8];  // per octant, indexed by instance
  std::vector<int32_t> shape_root;     // per shape: root node of its BLAS or -1
  int32_t top_root = -1;
  int64_t inlined_instances = 0, instanced_instances = 0, flattened_instances = 0;
};
// Returns 0 or a negative jt_status (message set).
int jt_build_wide(const std::vector<JtHostShape>& shapes, const std::vector<JtHostInstance>& instances,
                  const std::vector<jt_bvh_node>& tlas_nodes, const std::vector<int64_t>& tlas_prims,
                  JtWideResult* out);

// ---- staged scene (jt_stage.cpp): everything the device needs, still in host vectors ---------------
struct JtStagedScene {
  std::vector<JtHostShape> hshapes;
  std::vector<JtShapeRec> shape_recs;
  std::vector<float> positions, normals, texcoords;
  std::vector<float4> colors;
  std::vector<int4> elements;
  std::vector<float4> ref_nodes;
  std::vector<int32_t> ref_prims;
  std::vector<JtHostInstance> hinst;
  std::vector<JtInstanceRec> inst_recs;
  JtWideResult wide;
  JtBigVec<uint32_t> tri_rank;  // [8][num_wtris]
  std::vector<uint32_t> inst_rank;
  std::vector<float4> inst_bounds;
  std::vector<JtMaterialRec> mats;
  std::vector<JtTextureRec> texs;
  std::vector<float4> texels_f;
  std::vector<uchar4> texels_b;
  std::vector<JtEnvRec> envs;
  std::vector<JtLightRec> lights;
  std::vector<float> cdf;
  std::vector<int32_t> cdf_guide;
  std::vector<JtCameraRec> cams;
  std::vector<float> lut;
  int depth = 0, blas_depth = 0;
  bool wide_from_cache = false;
  int32_t tlas_num_nodes = 0, num_instances = 0, num_environments = 0, num_lights = 0, num_cameras = 0;
};
struct JtStagedPointers {
  const float4* ref_nodes = nullptr; const int32_t* ref_prims = nullptr; const JtShapeRec* shapes = nullptr;
  const float* positions = nullptr; const float* normals = nullptr; const float* texcoords = nullptr;
  const float4* colors = nullptr; const int4* elements = nullptr; const JtInstanceRec* instances = nullptr;
  const JtMaterialRec* materials = nullptr; const JtTextureRec* textures = nullptr;
  const float4* texels_f = nullptr; const uchar4* texels_b = nullptr; const float* srgb_lut = nullptr;
  const JtEnvRec* environments = nullptr; const JtLightRec* lights = nullptr; const float* light_cdf = nullptr;
  const int32_t* light_guide = nullptr;
  const JtCameraRec* cameras = nullptr; const float4* wnodes = nullptr; const float4* wtris = nullptr;
  const uint32_t* tri_rank = nullptr; const uint32_t* inst_rank = nullptr;
  const float4* inst_bounds = nullptr;
};
int jt_stage_scene(const jt_scene_desc* desc, JtStagedScene* out);
void jt_build_cdf_guide(const float* cdf, int64_t n, JtLightRec* rec, std::vector<int32_t>* guide);
void jt_fill_dev_scene(const JtStagedScene& S, const JtStagedPointers& P, JtDevScene* D);
