// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
//
// Scene containers of the oracle + restatement of the reference's BVH build and traversal:
//   /root/reference/src/bvh.jl:32-520, src/shape.jl:50-70, src/trace.jl:102-187 (lights)
// The scene arrives through the same jt_scene_desc the product consumes (include/jtrace_b200.h),
// but the oracle can ignore the BVH / lights given there and build its own from the shapes.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/jtrace_b200.h"
#include "orc_math.h"

namespace orc {

static const int64_t invalid_id = -1;  // src/scene.jl:45
static const int BVH_MAX_PRIMS = 4;    // src/bvh.jl:32

struct BvhNode {  // src/bvh.jl:34-39
  Bbox bbox;
  int64_t start;
  int16_t num;
  int8_t axis;
  bool internal;
};
inline BvhNode default_node() { return BvhNode{empty_bbox(), 0, 0, 1, false}; }  // :41

struct BvhTree {  // :46-49
  std::vector<BvhNode> nodes;
  std::vector<int64_t> primitives;  // 1-based
};

struct Shape {  // src/shape.jl:13-23
  std::vector<V3> positions, normals;
  std::vector<V2> texcoords;
  std::vector<V4> colors;
  std::vector<int64_t> triangles;  // 3 per element, 1-based
  std::vector<int64_t> quads;      // 4 per element, 1-based
  BvhTree bvh;
  int64_t ntri() const { return (int64_t)triangles.size() / 3; }
  int64_t nquad() const { return (int64_t)quads.size() / 4; }
};

struct Instance { Frame frame; int64_t shape, material; };
struct Texture {
  int64_t width, height;
  bool linear;
  std::vector<V4> pixelsf;
  std::vector<uint8_t> pixelsb;  // 4 per texel
};
struct Material {
  int32_t type;
  V3 emission, color;
  float roughness, metallic, ior;
  V3 scattering;
  float scanisotropy, trdepth, opacity;
  int64_t emission_tex, color_tex, roughness_tex, scattering_tex, normal_tex;
};
struct Environment { Frame frame; V3 emission; int64_t emission_tex; };
struct Camera { Frame frame; bool orthographic; float lens, film, aspect, focus, aperture; };
struct Light { int64_t instance, environment; std::vector<float> cdf; };

enum MaterialType { matte = 0, glossy, reflective, transparent, refractive, subsurface, volumetric, gltfpbr };

struct Scene {
  std::vector<Camera> cameras;
  std::vector<Instance> instances;
  std::vector<Environment> environments;
  std::vector<Shape> shapes;
  std::vector<Texture> textures;
  std::vector<Material> materials;
  std::vector<Light> lights;
  BvhTree bvh;  // TLAS
};

struct ShapeIsec { int64_t element; V2 uv; float distance; bool hit; };        // src/shape.jl:50-58
struct SceneIsec { int64_t instance, element; V2 uv; float distance; bool hit; };  // :61-70
inline ShapeIsec no_shape_isec() { return ShapeIsec{-1, V2{0, 0}, 0.0f, false}; }
inline SceneIsec no_scene_isec() { return SceneIsec{-1, -1, V2{0, 0}, 0.0f, false}; }

// Per-thread work counters (algorithmic bytes per ray)
struct Counters {
  uint64_t scene_rays = 0, light_rays = 0, camera_paths = 0;
  uint64_t tlas_nodes = 0, blas_nodes = 0, instance_visits = 0, tri_tests = 0, quad_tests = 0;
  // the share of the above spent inside intersect_instance_bvh (light-pdf probes)
  uint64_t probe_blas_nodes = 0, probe_tri_tests = 0, probe_quad_tests = 0;
  // optional log of every closest-hit query in call order (orc_trace_pixel_rays: divergence hunting in the tests)
  struct LoggedRay { Ray ray; int64_t instance; };  // instance = -1: intersect_scene_bvh, else intersect_instance_bvh
  std::vector<LoggedRay>* log = nullptr;
  void add(const Counters& o) {
    scene_rays += o.scene_rays; light_rays += o.light_rays; camera_paths += o.camera_paths;
    tlas_nodes += o.tlas_nodes; blas_nodes += o.blas_nodes; instance_visits += o.instance_visits;
    tri_tests += o.tri_tests; quad_tests += o.quad_tests;
    probe_blas_nodes += o.probe_blas_nodes; probe_tri_tests += o.probe_tri_tests; probe_quad_tests += o.probe_quad_tests;
  }
};

// ---- build: src/bvh.jl:138-304 ---------------------------------------------------------------
// partition, :281-304 (Hoare scheme on 1-based inclusive [start, stop]; returns j)
inline int64_t partition(const std::vector<V3>& centers, int axis, float split,
                         std::vector<int64_t>& prims, int64_t start, int64_t stop) {
  int64_t i = start, j = stop;
  while (true) {
    while (i <= stop && centers[prims[i - 1] - 1][axis - 1] < split) i++;
    while (j >= start && centers[prims[j - 1] - 1][axis - 1] >= split) j--;
    if (i >= j) break;
    std::swap(prims[i - 1], prims[j - 1]);
  }
  return j;
}

// split_middle, :185-216
inline void split_middle(std::vector<int64_t>& prims, const std::vector<Bbox>& /*bboxes*/,
                         const std::vector<V3>& centers, int64_t left, int64_t right,
                         int64_t* mid, int8_t* axis_out) {
  Bbox cb = empty_bbox();
  for (int64_t i = left; i <= right; i++) cb = merge(cb, centers[prims[i - 1] - 1]);
  V3 csize = cb.mx - cb.mn;
  if (csize == V3{0, 0, 0}) {
    *mid = (left + right + 1) / 2;
    *axis_out = 1;
    return;
  }
  int8_t axis = 1;
  if (csize.x >= csize.y && csize.x >= csize.z) axis = 1;
  if (csize.y >= csize.x && csize.y >= csize.z) axis = 2;
  if (csize.z >= csize.x && csize.z >= csize.y) axis = 3;
  float split = center(cb)[axis - 1];
  int64_t middle = partition(centers, axis, split, prims, left, right);
  if (middle < left || middle > right) {
    *mid = (left + right + 1) / 2;
    *axis_out = axis;
    return;
  }
  *mid = middle;
  *axis_out = axis;
}

// bbox_area, :276-279
inline float bbox_area(const Bbox& b) {
  V3 s = b.mx - b.mn;
  return ((0.000000000001f + (2.0f * s.x) * s.y) + (2.0f * s.x) * s.z) + (2.0f * s.y) * s.z;
}

// split_sah, :218-274
inline void split_sah(std::vector<int64_t>& prims, const std::vector<Bbox>& bboxes,
                      const std::vector<V3>& centers, int64_t left, int64_t right, int64_t* mid,
                      int8_t* axis_out) {
  Bbox cb = empty_bbox();
  for (int64_t i = left; i <= right; i++) cb = merge(cb, centers[prims[i - 1] - 1]);
  V3 csize = cb.mx - cb.mn;
  if (csize == V3{0, 0, 0}) {
    *mid = (left + right + 1) / 2;
    *axis_out = 1;
    return;
  }
  int8_t axis = 1;
  const int nbins = 16;
  float split = 0.0f;
  float min_cost = FLT_INF;  // typemax(Float32)
  for (int saxis = 1; saxis <= 3; saxis++) {
    for (int b = 1; b <= nbins - 1; b++) {
      float bsplit = cb.mn[saxis - 1] + ((float)b * csize[saxis - 1]) / (float)nbins;
      Bbox lb = empty_bbox(), rb = empty_bbox();
      int64_t ln = 0, rn = 0;
      for (int64_t i = left; i <= right; i++) {
        int64_t p = prims[i - 1] - 1;
        if (centers[p][saxis - 1] < bsplit) {
          lb = merge(lb, bboxes[p]);
          ln++;
        } else {
          rb = merge(rb, bboxes[p]);
          rn++;
        }
      }
      float cost = (1.0f + ((float)ln * bbox_area(lb)) / bbox_area(cb)) +
                   ((float)rn * bbox_area(rb)) / bbox_area(cb);
      if (cost < min_cost) {
        min_cost = cost;
        split = bsplit;
        axis = (int8_t)saxis;
      }
    }
  }
  int64_t middle = partition(centers, axis, split, prims, left, right);
  if (middle == left || middle == right) {
    *mid = (left + right + 1) / 2;
    *axis_out = axis;
    return;
  }
  *mid = middle;
  *axis_out = axis;
}

// make_bvh, :138-183
inline BvhTree make_bvh(const std::vector<Bbox>& bboxes, bool high_quality) {
  BvhTree bvh;
  int64_t n = (int64_t)bboxes.size();
  bvh.primitives.resize(n);
  for (int64_t i = 0; i < n; i++) bvh.primitives[i] = i + 1;
  std::vector<V3> centers(n);
  for (int64_t i = 0; i < n; i++) centers[i] = center(bboxes[i]);
  struct Item { int64_t node, left, right; };
  std::vector<Item> stack;
  stack.push_back(Item{1, 1, n});
  bvh.nodes.push_back(default_node());
  while (!stack.empty()) {
    Item it = stack.back();
    stack.pop_back();
    BvhNode node = bvh.nodes[it.node - 1];
    for (int64_t i = it.left; i <= it.right; i++)
      node.bbox = merge(node.bbox, bboxes[bvh.primitives[i - 1] - 1]);
    bvh.nodes[it.node - 1] = node;
    if (it.right - it.left + 1 > BVH_MAX_PRIMS) {
      int64_t mid;
      int8_t axis;
      if (high_quality)
        split_sah(bvh.primitives, bboxes, centers, it.left, it.right, &mid, &axis);
      else
        split_middle(bvh.primitives, bboxes, centers, it.left, it.right, &mid, &axis);
      int64_t start = (int64_t)bvh.nodes.size() + 1;
      bvh.nodes[it.node - 1] = BvhNode{node.bbox, start, 2, axis, true};
      bvh.nodes.push_back(default_node());
      bvh.nodes.push_back(default_node());
      stack.push_back(Item{start, it.left, mid});
      stack.push_back(Item{start + 1, mid + 1, it.right});
    } else {
      bvh.nodes[it.node - 1] =
          BvhNode{node.bbox, it.left, (int16_t)(it.right - it.left + 1), node.axis, false};
    }
  }
  return bvh;
}

// make_shape_bvh, :90-136 (triangles take precedence over quads)
inline BvhTree make_shape_bvh(const Shape& s, bool high_quality) {
  std::vector<Bbox> boxes;
  if (s.ntri() > 0) {
    boxes.resize(s.ntri());
    for (int64_t i = 0; i < s.ntri(); i++) {
      const int64_t* t = &s.triangles[3 * i];
      boxes[i] = triangle_bounds(s.positions[t[0] - 1], s.positions[t[1] - 1], s.positions[t[2] - 1]);
    }
  } else if (s.nquad() > 0) {
    boxes.resize(s.nquad());
    for (int64_t i = 0; i < s.nquad(); i++) {
      const int64_t* q = &s.quads[4 * i];
      boxes[i] = quad_bounds(s.positions[q[0] - 1], s.positions[q[1] - 1], s.positions[q[2] - 1],
                             s.positions[q[3] - 1]);
    }
  }
  return make_bvh(boxes, high_quality);
}

// make_scene_bvh, :66-88
inline void make_scene_bvh(Scene& scene, bool high_quality) {
  for (auto& s : scene.shapes) s.bvh = make_shape_bvh(s, high_quality);
  std::vector<Bbox> boxes(scene.instances.size());
  for (size_t i = 0; i < boxes.size(); i++) {
    const Instance& inst = scene.instances[i];
    const BvhTree& sb = scene.shapes[inst.shape - 1].bvh;
    boxes[i] = sb.nodes.empty() ? empty_bbox() : transform_bbox(inst.frame, sb.nodes[0].bbox);
  }
  scene.bvh = make_bvh(boxes, high_quality);
}

// ---- traversal: src/bvh.jl:306-520 -------------------------------------------------------------
// intersect_shape_bvh, :373-491 (points / lines unreachable: SURVEY.md §2.3)
inline ShapeIsec intersect_shape_bvh(const Shape& shape, Ray ray, bool find_any, Counters* cnt) {
  const BvhTree& bvh = shape.bvh;
  if (bvh.nodes.empty()) return no_shape_isec();
  int64_t stack[256];
  int node_cur = 0;
  stack[node_cur++] = 1;
  ShapeIsec isec = no_shape_isec();
  V3 dinv{1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z};
  int dsign[3] = {ray.d.x < 0 ? 1 : 0, ray.d.y < 0 ? 1 : 0, ray.d.z < 0 ? 1 : 0};
  const bool tris = shape.ntri() > 0;
  while (node_cur != 0) {
    const BvhNode& node = bvh.nodes[stack[--node_cur] - 1];
    if (cnt) cnt->blas_nodes++;
    if (!intersect_bbox(ray, dinv, node.bbox)) continue;
    if (node.internal) {
      // Q1: for d[axis] >= 0 push start then start+1 -> the FAR child is popped first
      if (dsign[node.axis - 1] == 0) {
        stack[node_cur++] = node.start;
        stack[node_cur++] = node.start + 1;
      } else {
        stack[node_cur++] = node.start + 1;
        stack[node_cur++] = node.start;
      }
    } else if (tris) {
      for (int64_t i = node.start; i <= node.start + node.num - 1; i++) {
        const int64_t* t = &shape.triangles[3 * (bvh.primitives[i - 1] - 1)];
        if (cnt) cnt->tri_tests++;
        PrimIsec p = intersect_triangle(ray, shape.positions[t[0] - 1], shape.positions[t[1] - 1],
                                        shape.positions[t[2] - 1]);
        if (!p.hit) continue;
        isec = ShapeIsec{bvh.primitives[i - 1], p.uv, p.distance, true};  // Q2: overwrite
        ray.tmax = p.distance;
      }
    } else if (shape.nquad() > 0) {
      for (int64_t i = node.start; i <= node.start + node.num - 1; i++) {
        const int64_t* q = &shape.quads[4 * (bvh.primitives[i - 1] - 1)];
        if (cnt) cnt->quad_tests++;
        PrimIsec p = intersect_quad(ray, shape.positions[q[0] - 1], shape.positions[q[1] - 1],
                                    shape.positions[q[2] - 1], shape.positions[q[3] - 1]);
        if (!p.hit) continue;
        isec = ShapeIsec{bvh.primitives[i - 1], p.uv, p.distance, true};
        ray.tmax = p.distance;
      }
    }
    if (find_any && isec.hit) return isec;
  }
  return isec;
}

// intersect_scene_bvh, :306-371 (Q5: inverse(frame, true) recomputed per visit)
inline SceneIsec intersect_scene_bvh(const Scene& scene, Ray ray, bool find_any, Counters* cnt) {
  const BvhTree& bvh = scene.bvh;
  if (cnt) cnt->scene_rays++;
  if (cnt && cnt->log) cnt->log->push_back(Counters::LoggedRay{ray, -1});
  if (bvh.nodes.empty()) return no_scene_isec();  // the reference throws here
  int64_t stack[256];
  int node_cur = 0;
  stack[node_cur++] = 1;
  SceneIsec isec = no_scene_isec();
  V3 dinv{1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z};
  int dsign[3] = {ray.d.x < 0 ? 1 : 0, ray.d.y < 0 ? 1 : 0, ray.d.z < 0 ? 1 : 0};
  while (node_cur != 0) {
    const BvhNode& node = bvh.nodes[stack[--node_cur] - 1];
    if (cnt) cnt->tlas_nodes++;
    if (!intersect_bbox(ray, dinv, node.bbox)) continue;
    if (node.internal) {
      if (dsign[node.axis - 1] == 0) {
        stack[node_cur++] = node.start;
        stack[node_cur++] = node.start + 1;
      } else {
        stack[node_cur++] = node.start + 1;
        stack[node_cur++] = node.start;
      }
    } else {
      for (int64_t i = node.start; i <= node.start + node.num - 1; i++) {
        const Instance& inst = scene.instances[bvh.primitives[i - 1] - 1];
        if (cnt) cnt->instance_visits++;
        Ray inv_ray = transform_ray(inverse(inst.frame, true), ray);
        ShapeIsec s = intersect_shape_bvh(scene.shapes[inst.shape - 1], inv_ray, find_any, cnt);
        if (!s.hit) continue;
        isec = SceneIsec{bvh.primitives[i - 1], s.element, s.uv, s.distance, true};
        ray.tmax = s.distance;
      }
    }
    if (find_any && isec.hit) return isec;
  }
  return isec;
}

// intersect_instance_bvh, :493-520
inline SceneIsec intersect_instance_bvh(const Scene& scene, int64_t instance_, Ray ray,
                                        bool find_any, Counters* cnt) {
  const Instance& inst = scene.instances[instance_ - 1];
  if (cnt) { cnt->light_rays++; cnt->instance_visits++; }
  if (cnt && cnt->log) cnt->log->push_back(Counters::LoggedRay{ray, instance_});
  Ray inv_ray = transform_ray(inverse(inst.frame, true), ray);
  uint64_t n0 = cnt ? cnt->blas_nodes : 0, t0 = cnt ? cnt->tri_tests : 0, q0 = cnt ? cnt->quad_tests : 0;
  ShapeIsec s = intersect_shape_bvh(scene.shapes[inst.shape - 1], inv_ray, find_any, cnt);
  if (cnt) {
    cnt->probe_blas_nodes += cnt->blas_nodes - n0;
    cnt->probe_tri_tests += cnt->tri_tests - t0;
    cnt->probe_quad_tests += cnt->quad_tests - q0;
  }
  if (!s.hit) return no_scene_isec();
  return SceneIsec{instance_, s.element, s.uv, s.distance, true};
}

}  // namespace orc
