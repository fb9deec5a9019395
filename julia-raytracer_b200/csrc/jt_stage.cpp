// jt_stage.cpp -- host-side staging of a jt_scene_desc: validates the description, converts the
// Julia-side conventions (1-based Int64 ids, 40-byte BvhNode, per-shape arrays) into the flat
// 0-based device layout of jt_internal.h and builds the wide BVH. Pure host code: the CUDA
// translation unit (jt_api.cu) only uploads what is staged here.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>

#include "jt_internal.h"

#ifndef JT_WIDE_STACK
#define JT_WIDE_STACK 64
#endif

static inline float4 mk_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline int4 mk_int4(int x, int y, int z, int w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#define make_float4 mk_float4
#define make_int4 mk_int4

static void inverse_frame(const float* f, float* out) {
  auto cross = [](const float* a, const float* b, float* r) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
  };
  const float *x = f, *y = f + 3, *z = f + 6, *o = f + 9;
  float c0[3], c1[3], c2[3];
  cross(y, z, c0);
  cross(z, x, c1);
  cross(x, y, c2);
  // adjoint = transpose(Mat3f(c0, c1, c2)): columns (c0.x,c1.x,c2.x), (c0.y,c1.y,c2.y), (c0.z,c1.z,c2.z)
  float det = (x[0] * c0[0] + x[1] * c0[1]) + x[2] * c0[2];
  float s = 1.0f / det;
  float m[9] = {c0[0] * s, c1[0] * s, c2[0] * s, c0[1] * s, c1[1] * s, c2[1] * s, c0[2] * s, c1[2] * s, c2[2] * s};
  for (int k = 0; k < 9; k++) out[k] = m[k];
  // -(minv * o) = -((m1*o1 + m2*o2) + m3*o3)
  for (int k = 0; k < 3; k++) out[9 + k] = -((m[k] * o[0] + m[3 + k] * o[1]) + m[6 + k] * o[2]);
}

static bool is_identity_frame(const float* f) {
  static const float id[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
  for (int k = 0; k < 12; k++)
    if (f[k] != id[k]) return false;
  return true;
}

static int tex_id(int64_t julia_id, int64_t count, const char* what, int32_t* out) {
  if (julia_id == -1 || julia_id == 0) {  // -1 = invalid_id; 0 never occurs (get(..., -2) + 1 = -1)
    *out = -1;
    return JT_OK;
  }
  if (julia_id < 1 || julia_id > count) return jt_set_error(JT_ERR_INVALID, "%s texture id %lld out of range", what, (long long)julia_id);
  *out = (int32_t)(julia_id - 1);
  return JT_OK;
}

static void pack_ref_nodes(const jt_bvh_node* nodes, int64_t n, int32_t child_off, int32_t prim_off,
                           std::vector<float4>* out) {
  for (int64_t i = 0; i < n; i++) {
    const jt_bvh_node& b = nodes[i];
    int32_t start = (int32_t)(b.start - 1) + (b.internal ? child_off : prim_off);
    int32_t packed = ((int32_t)b.num & 0xFFFF) | (((int32_t)(b.axis - 1) & 0xFF) << 16) | ((b.internal ? 1 : 0) << 24);
    float4 a = make_float4(b.bbox_min[0], b.bbox_min[1], b.bbox_min[2], b.bbox_max[0]);
    float4 c;
    c.x = b.bbox_max[1];
    c.y = b.bbox_max[2];
    memcpy(&c.z, &start, 4);
    memcpy(&c.w, &packed, 4);
    out->push_back(a);
    out->push_back(c);
  }
}

// Validate a host-built reference tree before it is walked by MODE_REF with its fixed int stack[JT_REF_STACK]
// (jt_dev_traverse.cuh): children and leaf ranges in range, every node reachable at most once from the root, and the
// traversal stack high-water mark (an internal node pops one entry and pushes two) within the stack. The reference
// raises BoundsError for the same inputs.
#define JT_REF_STACK_LIMIT 128 /* == JT_REF_STACK */
static int validate_ref_tree(const jt_bvh_node* nodes, int64_t n, int64_t num_prims, const char* what, long long id) {
  if (n <= 0) return JT_OK;
  std::vector<std::pair<int64_t, int>> st;  // node, stack entries below it while it is being visited
  st.push_back({0, 0});
  int64_t visited = 0;
  while (!st.empty()) {
    auto [i, below] = st.back();
    st.pop_back();
    if (++visited > n) return jt_set_error(JT_ERR_INVALID, "%s %lld: BVH is not a tree (a node is reachable twice)", what, id);
    const jt_bvh_node& b = nodes[i];
    if (b.internal) {
      if (b.axis < 1 || b.axis > 3) return jt_set_error(JT_ERR_INVALID, "%s %lld: BVH node %lld has axis %d", what, id, (long long)i, (int)b.axis);
      if (b.start < 1 || b.start + 1 > n) return jt_set_error(JT_ERR_INVALID, "%s %lld: BVH node %lld child index out of range", what, id, (long long)i);
      if (below + 2 > JT_REF_STACK_LIMIT)
        return jt_set_error(JT_ERR_UNSUPPORTED, "%s %lld: BVH too deep for the reference-order traversal stack (%d entries, --bvhstacksize default)", what, id, JT_REF_STACK_LIMIT);
      st.push_back({b.start - 1, below + 1});  // the child visited first still has its sibling below it
      st.push_back({b.start, below + 1});
    } else {
      if (b.num < 0 || b.start < 1 || b.start - 1 + b.num > num_prims)
        return jt_set_error(JT_ERR_INVALID, "%s %lld: BVH leaf %lld primitive range out of bounds", what, id, (long long)i);
    }
  }
  return JT_OK;
}

static int max_wide_depth(const JtBigVec<JtWideNode>& nodes, int root) {
  if (root < 0) return 0;
  int best = 0;
  std::vector<std::pair<int, int>> st;
  st.push_back({root, 1});
  while (!st.empty()) {
    auto [n, d] = st.back();
    st.pop_back();
    best = std::max(best, d);
    const JtWideNode& w = nodes[(size_t)n];
    int k = __builtin_popcount(w.imask);
    for (int i = 0; i < k; i++) st.push_back({(int)w.child_base + i, d + 1});
  }
  return best;
}


// Guide table for sample_discrete (src/sampling.jl:33-56) on a long CDF: the 17-21 dependent loads of the
// binary search are the hottest line of the shade kernel (profiles/r01/hot_lines_shade_v3.txt).
// g(x) = clamp((int)(x * scale), 0, K - 1) is monotone in x, so for a search key `limit` in bucket b every
// entry with g(c_i) < b is <= limit and every entry with g(c_i) > b is > limit: upper_bound's answer
// 1 + #{c_i <= limit} lies in [guide[b] + 1, guide[b + 1] + 1] with guide[b] = #{i : g(c_i) < b}. The device runs
// the reference's own bisection on that bracket, so the returned index is identical. Requires a
// non-decreasing CDF (sequential sums of non-negative weights are); otherwise no table is built.
void jt_build_cdf_guide(const float* c, int64_t n, JtLightRec* R, std::vector<int32_t>* guide) {
  R->guide_off = 0;
  R->guide_len = 0;
  R->guide_scale = 0.0f;
  R->_pad = 0;
  if (n < 64) return;
  float last = c[n - 1];
  if (!(last > 0.0f) || !std::isfinite(last)) return;
  for (int64_t i = 1; i < n; i++)
    if (!(c[i] >= c[i - 1])) return;
  if (!(c[0] >= 0.0f)) return;
  const int K = (int)std::min<int64_t>(n / 4, 1 << 20);
  const float scale = (float)K / last;
  if (!std::isfinite(scale) || !std::isfinite(scale * last)) return;
  auto g = [&](float x) {
    int b = (int)(x * scale);
    return std::max(0, std::min(K - 1, b));
  };
  R->guide_off = (int32_t)guide->size();
  R->guide_len = K;
  R->guide_scale = scale;
  guide->resize(guide->size() + (size_t)K + 1, 0);
  int32_t* G = guide->data() + R->guide_off;
  // histogram of g(c_i), then exclusive prefix sums: G[b] = #{i : g(c_i) < b}
  for (int64_t i = 0; i < n; i++) G[g(c[i]) + 1]++;
  for (int b = 1; b <= K; b++) G[b] += G[b - 1];
}

// ---- wide-BVH cache -------------------------------------------------------------------------------------------------
// File = header {magic, version, key, counts} + the raw arrays of JtWideResult. The key hashes every input of
// jt_build_wide and the builder's tuning knobs, so a stale or foreign file is never used; a short or damaged file fails
// the size check and is rebuilt.
static std::string g_bvh_cache_dir;
static bool g_bvh_cache_dir_set = false;
extern "C" JT_API int jt_set_bvh_cache_dir(const char* dir) {
  g_bvh_cache_dir = dir ? dir : "";
  g_bvh_cache_dir_set = true;
  return JT_OK;
}
static std::string bvh_cache_dir() {
  if (g_bvh_cache_dir_set) return g_bvh_cache_dir;
  const char* e = getenv("JT_BVH_CACHE_DIR");
  return e ? e : "";
}
struct WideHash {
  uint64_t h = 0x9E3779B97F4A7C15ull;
  void word(uint64_t w) {
    h ^= w;
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
  }
  void bytes(const void* p, size_t n) {
    const uint8_t* b = (const uint8_t*)p;
    word((uint64_t)n);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
      uint64_t w;
      memcpy(&w, b + i, 8);
      word(w);
    }
    uint64_t tail = 0;
    if (i < n) memcpy(&tail, b + i, n - i);
    word(tail);
  }
  template <class T>
  void vec(const std::vector<T>& v) { bytes(v.data(), v.size() * sizeof(T)); }
};
static const uint32_t kWideCacheVersion = 5;  // bump when the builder or the record layouts change
static uint64_t wide_cache_key(const std::vector<JtHostShape>& shapes, const std::vector<JtHostInstance>& inst,
                               const std::vector<jt_bvh_node>& tlas_nodes, const std::vector<int64_t>& tlas_prims) {
  if (bvh_cache_dir().empty()) return 0;
  WideHash H;
  H.word(kWideCacheVersion);
  for (const char* knob : {"JT_BRAID_MAX", "JT_BRAID_MIN_INSTANCES", "JT_LEAF_MAX", "JT_TRI_COST", "JT_BUILD_PARALLEL_MIN", "JT_COLLAPSE", "JT_COLLAPSE_DP_MAX_NODES"}) {
    const char* e = getenv(knob);
    H.bytes(e ? e : "", e ? strlen(e) : 0);
  }
  H.word(shapes.size());
  for (const JtHostShape& s : shapes) {
    H.word((uint64_t)s.kind);
    H.vec(s.pos);
    H.vec(s.elems);
    H.word(s.ref_nodes.size());
    for (const jt_bvh_node& n : s.ref_nodes) {  // field by field: the records carry padding
      H.bytes(n.bbox_min, 24);
      H.word((uint64_t)n.start);
      H.word(((uint64_t)(uint16_t)n.num << 16) | ((uint64_t)(uint8_t)n.axis << 8) | (uint64_t)n.internal);
    }
    H.vec(s.ref_prims);
  }
  H.word(inst.size());
  for (const JtHostInstance& i : inst) {
    H.bytes(i.frame, 48);
    H.word(((uint64_t)(uint32_t)i.shape << 32) | (uint64_t)(i.inlined ? 1 : 0));
  }
  H.word(tlas_nodes.size());
  for (const jt_bvh_node& n : tlas_nodes) {
    H.bytes(n.bbox_min, 24);
    H.word((uint64_t)n.start);
    H.word(((uint64_t)(uint16_t)n.num << 16) | ((uint64_t)(uint8_t)n.axis << 8) | (uint64_t)n.internal);
  }
  H.vec(tlas_prims);
  return H.h | 1ull;  // never 0 (= caching off)
}
struct WideCacheHeader {
  char magic[8];
  uint32_t version, reserved;
  uint64_t key;
  uint64_t nodes, tris, instances, shapes;
  int64_t top_root, inlined, instanced, flattened, depth, blas_depth;
};
static std::string wide_cache_path(uint64_t key) {
  char name[64];
  snprintf(name, sizeof(name), "/jtwide_%016llx.bin", (unsigned long long)key);
  return bvh_cache_dir() + name;
}
// Sections after the header, in file order: nodes, triangle records, tri_rank[8][tris], inst_rank[8][instances], shape_root.
// Loaded with parallel preads straight into the staged scene's (uninitialised) vectors.
static bool wide_cache_load(uint64_t key, JtStagedScene* S) {
  if (key == 0) return false;
  const int fd = open(wide_cache_path(key).c_str(), O_RDONLY);
  if (fd < 0) return false;
  WideCacheHeader h;
  bool ok = pread(fd, &h, sizeof(h), 0) == (ssize_t)sizeof(h) && !memcmp(h.magic, "JTWIDE\0", 8) &&
            h.version == kWideCacheVersion && h.key == key;
  if (ok) {
    const uint64_t expect = sizeof(h) + h.nodes * sizeof(JtWideNode) + h.tris * (sizeof(JtWideTri) + 32) + h.instances * 32 + h.shapes * 4;
    struct stat st;
    ok = fstat(fd, &st) == 0 && (uint64_t)st.st_size == expect;
  }
  if (ok) {
    JtWideResult& W = S->wide;
    W.nodes.resize(h.nodes);
    W.tris.resize(h.tris);
    S->tri_rank.resize(8 * h.tris);
    S->inst_rank.resize(8 * h.instances);
    W.shape_root.resize(h.shapes);
    struct Section { char* dst; uint64_t bytes; };
    const Section sections[5] = {{(char*)W.nodes.data(), h.nodes * sizeof(JtWideNode)}, {(char*)W.tris.data(), h.tris * sizeof(JtWideTri)},
                                 {(char*)S->tri_rank.data(), 8 * h.tris * 4}, {(char*)S->inst_rank.data(), 8 * h.instances * 4},
                                 {(char*)W.shape_root.data(), h.shapes * 4}};
    struct Piece { char* dst; uint64_t off, bytes; };
    std::vector<Piece> pieces;
    uint64_t off = sizeof(h);
    const uint64_t chunk = 32ull << 20;
    for (const Section& sec : sections) {
      for (uint64_t o = 0; o < sec.bytes; o += chunk) pieces.push_back({sec.dst + o, off + o, std::min(chunk, sec.bytes - o)});
      off += sec.bytes;
    }
    int failed = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : failed)
    for (int64_t i = 0; i < (int64_t)pieces.size(); i++) {
      uint64_t got = 0;
      while (got < pieces[(size_t)i].bytes) {
        ssize_t r = pread(fd, pieces[(size_t)i].dst + got, pieces[(size_t)i].bytes - got, (off_t)(pieces[(size_t)i].off + got));
        if (r <= 0) { failed++; break; }
        got += (uint64_t)r;
      }
    }
    ok = failed == 0;
    if (ok) {
      W.top_root = (int32_t)h.top_root;
      W.inlined_instances = h.inlined; W.instanced_instances = h.instanced; W.flattened_instances = h.flattened;
      S->depth = (int)h.depth;
      S->blas_depth = (int)h.blas_depth;
    } else {
      W = JtWideResult();
      S->tri_rank.clear();
      S->inst_rank.clear();
    }
  }
  close(fd);
  return ok;
}
static void wide_cache_store(uint64_t key, const JtStagedScene& S) {
  if (key == 0) return;
  const JtWideResult& W = S.wide;
  if (S.tri_rank.size() != 8 * W.tris.size() || S.inst_rank.size() % 8 != 0) return;
  const std::string path = wide_cache_path(key), tmp = path + ".tmp" + std::to_string((long long)getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;  // an unwritable cache directory only costs the next build
  WideCacheHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "JTWIDE\0", 8);
  h.version = kWideCacheVersion;
  h.key = key;
  h.nodes = W.nodes.size(); h.tris = W.tris.size(); h.instances = S.inst_rank.size() / 8; h.shapes = W.shape_root.size();
  h.top_root = W.top_root; h.inlined = W.inlined_instances; h.instanced = W.instanced_instances; h.flattened = W.flattened_instances;
  h.depth = S.depth; h.blas_depth = S.blas_depth;
  auto put = [&](const void* p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; };
  bool ok = put(&h, sizeof(h)) && put(W.nodes.data(), W.nodes.size() * sizeof(JtWideNode)) &&
            put(W.tris.data(), W.tris.size() * sizeof(JtWideTri)) && put(S.tri_rank.data(), S.tri_rank.size() * 4) &&
            put(S.inst_rank.data(), S.inst_rank.size() * 4) && put(W.shape_root.data(), W.shape_root.size() * 4);
  ok = (fclose(f) == 0) && ok;
  if (ok) ok = rename(tmp.c_str(), path.c_str()) == 0;  // atomic: concurrent ranks never see a partial file
  if (!ok) remove(tmp.c_str());
}

int jt_stage_scene(const jt_scene_desc* d, JtStagedScene* S) {
  auto& hshapes = S->hshapes; auto& shape_recs = S->shape_recs; auto& positions = S->positions;
  auto& normals = S->normals; auto& texcoords = S->texcoords; auto& colors = S->colors;
  auto& elements = S->elements; auto& ref_nodes = S->ref_nodes; auto& ref_prims = S->ref_prims;
  auto& hinst = S->hinst; auto& inst_recs = S->inst_recs; auto& wide = S->wide; auto& tri_rank = S->tri_rank;
  auto& inst_rank = S->inst_rank; auto& mats = S->mats; auto& texs = S->texs; auto& texels_f = S->texels_f;
  auto& texels_b = S->texels_b; auto& envs = S->envs; auto& lights = S->lights; auto& cdf = S->cdf;
  auto& cams = S->cams; auto& lut = S->lut; int& depth = S->depth; int& blas_depth = S->blas_depth;
  int rc = JT_OK;
  const bool verbose = getenv("JT_STAGE_VERBOSE") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    auto now = std::chrono::steady_clock::now();
    if (verbose) fprintf(stderr, "jt_stage_scene: %-28s %8.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
    t_last = now;
  };
  if (!d) return jt_set_error(JT_ERR_INVALID, "null scene description");
  if (d->num_cameras < 1 || !d->cameras) return jt_set_error(JT_ERR_INVALID, "scene has no camera");
  if (d->num_instances < 0 || d->num_shapes < 0 || d->num_materials < 0 || d->num_textures < 0 ||
      d->num_environments < 0 || d->num_lights < 0)
    return jt_set_error(JT_ERR_INVALID, "negative count in scene description");
  if (d->num_instances > 0 && (d->bvh.num_nodes <= 0 || !d->bvh.nodes || !d->bvh.primitives))
    return jt_set_error(JT_ERR_INVALID, "scene BVH (SceneBvh.bvh) missing");
  hshapes.assign((size_t)d->num_shapes, JtHostShape());
  shape_recs.assign((size_t)d->num_shapes, JtShapeRec());
  hinst.assign((size_t)d->num_instances, JtHostInstance());
  inst_recs.assign((size_t)d->num_instances, JtInstanceRec());
  mats.assign((size_t)d->num_materials, JtMaterialRec());
  texs.assign((size_t)d->num_textures, JtTextureRec());
  envs.assign((size_t)d->num_environments, JtEnvRec());
  lights.assign((size_t)d->num_lights, JtLightRec());
  cams.assign((size_t)d->num_cameras, JtCameraRec());
  lut.assign(256, 0.0f);

  // ---- shapes: geometry arrays + reference BVHs --------------------------------------------------
  // TLAS first
  if (int vrc = validate_ref_tree(d->bvh.nodes, d->bvh.num_nodes, d->bvh.num_primitives, "scene", 0)) return vrc;
  pack_ref_nodes(d->bvh.nodes, d->bvh.num_nodes, 0, 0, &ref_nodes);
  for (int64_t i = 0; i < d->bvh.num_primitives; i++) {
    int64_t p = d->bvh.primitives[i];
    if (p < 1 || p > d->num_instances) return jt_set_error(JT_ERR_INVALID, "scene BVH instance id out of range");
    ref_prims.push_back((int32_t)(p - 1));
  }
  for (int64_t s = 0; s < d->num_shapes; s++) {
    const jt_shape_desc& h = d->shapes[s];
    JtHostShape& H = hshapes[(size_t)s];
    JtShapeRec& R = shape_recs[(size_t)s];
    memset(&R, 0, sizeof(R));
    R.norm_off = R.uv_off = R.col_off = -1;
    R.wide_root = -1;
    int64_t nv = h.num_positions;
    if (h.num_triangles > 0) H.kind = 1;
    else if (h.num_quads > 0) H.kind = 2;
    int64_t ne = H.kind == 1 ? h.num_triangles : (H.kind == 2 ? h.num_quads : 0);
    R.kind = H.kind;
    R.num_elements = (int32_t)ne;
    R.pos_off = (int32_t)(positions.size() / 3);
    R.elem_off = (int32_t)elements.size();
    if (nv > 0 && !h.positions) return jt_set_error(JT_ERR_INVALID, "shape %lld: positions missing", (long long)s);
    H.pos.assign(h.positions, h.positions + 3 * nv);
    positions.insert(positions.end(), H.pos.begin(), H.pos.end());
    if (h.num_normals > 0) {
      if (h.num_normals != nv) return jt_set_error(JT_ERR_INVALID, "shape %lld: normals/positions size mismatch", (long long)s);
      R.norm_off = (int32_t)(normals.size() / 3);
      normals.insert(normals.end(), h.normals, h.normals + 3 * nv);
    }
    if (h.num_texcoords > 0) {
      if (h.num_texcoords != nv) return jt_set_error(JT_ERR_INVALID, "shape %lld: texcoords/positions size mismatch", (long long)s);
      R.uv_off = (int32_t)(texcoords.size() / 2);
      texcoords.insert(texcoords.end(), h.texcoords, h.texcoords + 2 * nv);
    }
    if (h.num_colors > 0) {
      if (h.num_colors != nv) return jt_set_error(JT_ERR_INVALID, "shape %lld: colors/positions size mismatch", (long long)s);
      R.col_off = (int32_t)colors.size();
      for (int64_t v = 0; v < nv; v++)
        colors.push_back(make_float4(h.colors[4 * v], h.colors[4 * v + 1], h.colors[4 * v + 2], h.colors[4 * v + 3]));
    }
    H.elems.resize(4 * (size_t)ne);
    for (int64_t e = 0; e < ne; e++) {
      int64_t v[4];
      if (H.kind == 1) {
        v[0] = h.triangles[3 * e]; v[1] = h.triangles[3 * e + 1]; v[2] = h.triangles[3 * e + 2]; v[3] = v[2];
      } else {
        v[0] = h.quads[4 * e]; v[1] = h.quads[4 * e + 1]; v[2] = h.quads[4 * e + 2]; v[3] = h.quads[4 * e + 3];
      }
      for (int k = 0; k < 4; k++) {
        if (v[k] < 1 || v[k] > nv) return jt_set_error(JT_ERR_INVALID, "shape %lld element %lld: vertex id out of range", (long long)s, (long long)e);
        H.elems[4 * (size_t)e + k] = (int32_t)(v[k] - 1);
      }
      elements.push_back(make_int4(H.elems[4 * e], H.elems[4 * e + 1], H.elems[4 * e + 2], H.elems[4 * e + 3]));
    }
    if (ne > 0) {
      if (h.bvh.num_nodes <= 0 || !h.bvh.nodes || !h.bvh.primitives || h.bvh.num_primitives != ne)
        return jt_set_error(JT_ERR_INVALID, "shape %lld: ShapeBvh missing or inconsistent", (long long)s);
      H.ref_nodes.assign(h.bvh.nodes, h.bvh.nodes + h.bvh.num_nodes);
      H.ref_prims.assign(h.bvh.primitives, h.bvh.primitives + h.bvh.num_primitives);
      R.ref_node_off = (int32_t)(ref_nodes.size() / 2);
      R.ref_prim_off = (int32_t)ref_prims.size();
      R.num_ref_nodes = (int32_t)h.bvh.num_nodes;
      if (int vrc = validate_ref_tree(h.bvh.nodes, h.bvh.num_nodes, ne, "shape", (long long)s)) return vrc;
      pack_ref_nodes(h.bvh.nodes, h.bvh.num_nodes, 0, 0, &ref_nodes);  // shape-local indices
      for (int64_t i = 0; i < ne; i++) {
        int64_t p = h.bvh.primitives[i];
        if (p < 1 || p > ne) return jt_set_error(JT_ERR_INVALID, "shape %lld: BVH primitive id out of range", (long long)s);
        ref_prims.push_back((int32_t)(p - 1));
      }
    }
  }

  lap("shapes + reference trees");
  // ---- instances ------------------------------------------------------------------------------------
  for (int64_t i = 0; i < d->num_instances; i++) {
    const jt_instance& in = d->instances[i];
    if (in.shape < 1 || in.shape > d->num_shapes) return jt_set_error(JT_ERR_INVALID, "instance %lld: shape id out of range", (long long)i);
    if (in.material < 1 || in.material > d->num_materials) return jt_set_error(JT_ERR_INVALID, "instance %lld: material id out of range", (long long)i);
    JtHostInstance& H = hinst[(size_t)i];
    memcpy(H.frame, &in.frame, 48);
    inverse_frame(H.frame, H.inv);
    H.shape = (int)in.shape - 1;
    H.material = (int)in.material - 1;
    H.inlined = is_identity_frame(H.frame);
    JtInstanceRec& R = inst_recs[(size_t)i];
    memcpy(R.frame, H.frame, 48);
    memcpy(R.inv, H.inv, 48);
    R.shape = H.shape;
    R.material = H.material;
    R.inlined = H.inlined ? 1 : 0;
    R._pad = 0;
  }

  // ---- padded world-space box per instance: lets the light-pdf probes skip instances the ray cannot reach
  S->inst_bounds.assign(2 * (size_t)d->num_instances, make_float4(0, 0, 0, 0));
  for (int64_t i = 0; i < d->num_instances; i++) {
    const JtHostInstance& H = hinst[(size_t)i];
    const JtHostShape& sh = hshapes[(size_t)H.shape];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (!sh.ref_nodes.empty()) {
      const jt_bvh_node& rn = sh.ref_nodes[0];
      for (int c = 0; c < 8; c++) {
        float p[3] = {(c & 4) ? rn.bbox_max[0] : rn.bbox_min[0], (c & 2) ? rn.bbox_max[1] : rn.bbox_min[1],
                      (c & 1) ? rn.bbox_max[2] : rn.bbox_min[2]};
        for (int k = 0; k < 3; k++) {
          float w = ((H.frame[k] * p[0] + H.frame[3 + k] * p[1]) + H.frame[6 + k] * p[2]) + H.frame[9 + k];
          lo[k] = std::fmin(lo[k], w);
          hi[k] = std::fmax(hi[k], w);
        }
      }
      for (int k = 0; k < 3; k++) {  // generous slack: the probe itself is tested in instance space
        float pad = 1e-4f * std::fmax(std::fabs(lo[k]), std::fabs(hi[k])) + 1e-5f * (hi[k] - lo[k]) + 1e-30f;
        lo[k] -= pad;
        hi[k] += pad;
      }
    }
    S->inst_bounds[2 * (size_t)i] = make_float4(lo[0], lo[1], lo[2], 0.0f);
    S->inst_bounds[2 * (size_t)i + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
  }

  // ---- wide BVH ----------------------------------------------------------------------------------------
  std::vector<jt_bvh_node> tlas_nodes(d->bvh.nodes, d->bvh.nodes + d->bvh.num_nodes);
  std::vector<int64_t> tlas_prims(d->bvh.primitives, d->bvh.primitives + d->bvh.num_primitives);
  // N1: the build is the expensive host step for instancing-heavy scenes (ecosys: 16.8 M flattened
  // records, 4.7 s on 16 cores). With a cache directory set (jt_set_bvh_cache_dir / JT_BVH_CACHE_DIR) the finished wide
  // BVH is stored under a hash of everything the builder reads and found again by the next jt_scene_create.
  lap("instances");
  const uint64_t cache_key = wide_cache_key(hshapes, hinst, tlas_nodes, tlas_prims);
  lap("cache key");
  S->wide_from_cache = wide_cache_load(cache_key, S);
  lap("cache lookup");
  if (!S->wide_from_cache) {
    rc = jt_build_wide(hshapes, hinst, tlas_nodes, tlas_prims, &wide);
    if (rc != JT_OK) return rc;
    lap("wide BVH build");
    depth = max_wide_depth(wide.nodes, wide.top_root);
    blas_depth = 0;
    for (size_t s = 0; s < hshapes.size(); s++) blas_depth = std::max(blas_depth, max_wide_depth(wide.nodes, wide.shape_root[s]));
    for (int o = 0; o < 8; o++) {
      tri_rank.insert(tri_rank.end(), wide.tri_rank[o].begin(), wide.tri_rank[o].end());
      JtBigVec<uint32_t>().swap(wide.tri_rank[o]);  // the concatenated table is what the device reads
    }
    for (int o = 0; o < 8; o++) inst_rank.insert(inst_rank.end(), wide.inst_rank[o].begin(), wide.inst_rank[o].end());
  }
  for (size_t s = 0; s < hshapes.size(); s++) shape_recs[s].wide_root = wide.shape_root[s];
  // per level at most one deferred node group + one postponed triangle group (jt_dev_persist.cuh)
  if (2 * (depth + blas_depth) + 4 > JT_WIDE_STACK)
    return jt_set_error(JT_ERR_UNSUPPORTED, "wide BVH too deep for the traversal stack (%d + %d levels, limit %d)",
                        depth, blas_depth, JT_WIDE_STACK - 4);
  if (!S->wide_from_cache) {
    wide_cache_store(cache_key, *S);
    lap("cache store");
  }
  lap("depth check + rank tables");

  // ---- materials, textures, environments, lights, cameras ------------------------------------------------
  for (int64_t i = 0; i < d->num_materials; i++) {
    const jt_material& m = d->materials[i];
    JtMaterialRec& R = mats[(size_t)i];
    memset(&R, 0, sizeof(R));
    if (m.type < 0 || m.type > 7) return jt_set_error(JT_ERR_INVALID, "material %lld: bad type %d", (long long)i, m.type);
    if (m.type == 7)
      return jt_set_error(JT_ERR_UNSUPPORTED, "material %lld: gltfpbr throws UndefVarError in the reference (src/trace.jl:743)", (long long)i);
    R.type = m.type;
    memcpy(R.emission, m.emission, 12);
    memcpy(R.color, m.color, 12);
    R.roughness = m.roughness; R.metallic = m.metallic; R.ior = m.ior;
    memcpy(R.scattering, m.scattering, 12);
    R.scanisotropy = m.scanisotropy; R.trdepth = m.trdepth; R.opacity = m.opacity;
    if ((rc = tex_id(m.emission_tex, d->num_textures, "emission", &R.emission_tex))) return rc;
    if ((rc = tex_id(m.color_tex, d->num_textures, "color", &R.color_tex))) return rc;
    if ((rc = tex_id(m.roughness_tex, d->num_textures, "roughness", &R.roughness_tex))) return rc;
    if ((rc = tex_id(m.scattering_tex, d->num_textures, "scattering", &R.scattering_tex))) return rc;
    if ((rc = tex_id(m.normal_tex, d->num_textures, "normal", &R.normal_tex))) return rc;
  }
  for (int64_t i = 0; i < d->num_textures; i++) {
    const jt_texture_desc& t = d->textures[i];
    JtTextureRec& R = texs[(size_t)i];
    memset(&R, 0, sizeof(R));
    if (t.width < 0 || t.height < 0 || t.width > 65536 || t.height > 65536) return jt_set_error(JT_ERR_INVALID, "texture %lld: bad size", (long long)i);
    R.width = (int32_t)t.width; R.height = (int32_t)t.height; R.linear = t.linear ? 1 : 0;
    int64_t n = t.width * t.height;
    if (t.pixelsf) {
      R.is_float = 1;
      R.offset = (int64_t)texels_f.size();
      const float4* src = (const float4*)t.pixelsf;
      texels_f.insert(texels_f.end(), src, src + n);
    } else if (t.pixelsb) {
      R.is_float = 0;
      R.offset = (int64_t)texels_b.size();
      const uchar4* src = (const uchar4*)t.pixelsb;
      texels_b.insert(texels_b.end(), src, src + n);
    } else if (n > 0) {
      return jt_set_error(JT_ERR_INVALID, "texture %lld: no pixel data", (long long)i);
    }
  }
  for (int64_t i = 0; i < d->num_environments; i++) {
    const jt_environment& e = d->environments[i];
    memcpy(envs[(size_t)i].frame, &e.frame, 48);
    memcpy(envs[(size_t)i].emission, e.emission, 12);
    if ((rc = tex_id(e.emission_tex, d->num_textures, "environment", &envs[(size_t)i].emission_tex))) return rc;
  }
  for (int64_t i = 0; i < d->num_lights; i++) {
    const jt_light_desc& l = d->lights[i];
    JtLightRec& R = lights[(size_t)i];
    R.instance = l.instance >= 1 ? (int32_t)(l.instance - 1) : -1;
    R.environment = l.environment >= 1 ? (int32_t)(l.environment - 1) : -1;
    if (R.instance >= d->num_instances || R.environment >= d->num_environments)
      return jt_set_error(JT_ERR_INVALID, "light %lld: id out of range", (long long)i);
    if (R.instance < 0 && R.environment >= 0 && envs[(size_t)R.environment].emission_tex < 0)
      return jt_set_error(JT_ERR_UNSUPPORTED, "light %lld: emissive environment without texture calls the undefined sample_sphere in the reference (src/trace.jl:1003)", (long long)i);
    if (l.num_elements <= 0 || !l.elements_cdf) return jt_set_error(JT_ERR_INVALID, "light %lld: empty elements_cdf", (long long)i);
    if (R.instance >= 0 && l.num_elements != shape_recs[(size_t)hinst[(size_t)R.instance].shape].num_elements)
      return jt_set_error(JT_ERR_INVALID, "light %lld: elements_cdf length does not match the shape", (long long)i);
    if (R.instance < 0 && R.environment >= 0) {
      const JtTextureRec& T = texs[(size_t)envs[(size_t)R.environment].emission_tex];
      if (l.num_elements != (int64_t)T.width * T.height) return jt_set_error(JT_ERR_INVALID, "light %lld: elements_cdf length does not match the environment texture", (long long)i);
    }
    R.cdf_off = (int32_t)cdf.size();
    R.cdf_len = (int32_t)l.num_elements;
    cdf.insert(cdf.end(), l.elements_cdf, l.elements_cdf + l.num_elements);
    jt_build_cdf_guide(l.elements_cdf, l.num_elements, &R, &S->cdf_guide);
  }
  for (int64_t i = 0; i < d->num_cameras; i++) {
    const jt_camera& c = d->cameras[i];
    JtCameraRec& R = cams[(size_t)i];
    memset(&R, 0, sizeof(R));
    memcpy(R.frame, &c.frame, 48);
    R.orthographic = c.orthographic; R.lens = c.lens; R.film = c.film; R.aspect = c.aspect; R.focus = c.focus;
    R.aperture = c.aperture;
  }
  if (d->srgb_to_rgb_lut) {
    memcpy(lut.data(), d->srgb_to_rgb_lut, 1024);
  } else {
    for (int b = 0; b < 256; b++) {
      float c = (float)b / 255.0f;
      lut[(size_t)b] = c <= 0.04045f ? c / 12.92f : (float)exp2(log2((double)((c + 0.055f) / 1.055f)) * (double)2.4f);
    }
  }

  S->tlas_num_nodes = (int32_t)d->bvh.num_nodes;
  S->num_instances = (int32_t)d->num_instances;
  S->num_environments = (int32_t)d->num_environments;
  S->num_lights = (int32_t)d->num_lights;
  S->num_cameras = (int32_t)d->num_cameras;
  lap("materials, textures, lights");
  return JT_OK;
}

// Point a JtDevScene at staged arrays (host pointers for the CPU emulation used by the tests, or
// the device copies made by jt_api.cu).
void jt_fill_dev_scene(const JtStagedScene& S, const JtStagedPointers& P, JtDevScene* D) {
  memset(D, 0, sizeof(*D));
  D->ref_nodes = P.ref_nodes; D->ref_prims = P.ref_prims; D->shapes = P.shapes; D->positions = P.positions;
  D->normals = P.normals; D->texcoords = P.texcoords; D->colors = P.colors; D->elements = P.elements;
  D->instances = P.instances; D->materials = P.materials; D->textures = P.textures; D->texels_f = P.texels_f;
  D->texels_b = P.texels_b; D->srgb_lut = P.srgb_lut; D->environments = P.environments; D->lights = P.lights;
  D->light_cdf = P.light_cdf; D->light_guide = P.light_guide; D->cameras = P.cameras; D->wnodes = P.wnodes; D->wtris = P.wtris;
  D->tri_rank = P.tri_rank; D->inst_rank = P.inst_rank; D->inst_bounds = P.inst_bounds;
  D->tlas_num_nodes = S.tlas_num_nodes; D->num_instances = S.num_instances;
  D->num_environments = S.num_environments; D->num_lights = S.num_lights;
  D->num_wtris = (int32_t)S.wide.tris.size(); D->wide_root = S.wide.top_root;
}
