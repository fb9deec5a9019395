// jt_dev_traverse.cuh -- closest-hit queries.
//
// Replaces intersect_scene_bvh (src/bvh.jl:306-371), intersect_shape_bvh (:373-491) and
// intersect_instance_bvh (:493-520), with intersect_bbox / intersect_triangle / intersect_quad
// (src/geometry.jl:96-105, :206-258) as leaf and node tests.
//
// Two implementations behind one interface (template parameter MODE):
//   MODE_REF  -- the host-built binary BVH walked in the reference's own order (far child first,
//                Q1), `t <= tmax` acceptance with overwrite (Q2), Float64 slab fudge (Q3),
//                NaN-propagating min/max (Q4). Bit-exact by construction; it is the parity mode.
//   MODE_WIDE -- the 8-wide quantised BVH of jt_wide_bvh.cpp: octant-ordered near-first descent,
//                128-bit node/triangle fetches, one traversal for both levels (instance records
//                switch the ray into instance space and back). The leaf test is the same exact
//                arithmetic; exact-t ties fall back to the reference's visit-rank tables.
#pragma once
#include "jt_dev_math.cuh"

enum { MODE_WIDE = 0, MODE_REF = 1 };

// work counters, compiled in only by the host emulation of tests/emu (never in the CUDA build)
#ifdef JT_EMU_COUNT
#define JT_COUNT(name) (++jt_emu_counts.name)
#else
#define JT_COUNT(name) ((void)0)
#endif

// ---- exact leaf test --------------------------------------------------------------------------
// intersect_triangle (src/geometry.jl:206-236) on a record that stores p1, p2-p1, p3-p1.
// Returns true and (t,u,v) when the reference's test accepts against [tmin, tmax].
JT_DEV bool tri_test(f3 o, f3 d, float tmin, float tmax, f3 p1, f3 edge1, f3 edge2, float* t, float* u,
                     float* v) {
  f3 pvec = cross3(d, edge2);
  float det = dot3(edge1, pvec);
  if (det == 0.0f) return false;
  float inv_det = 1.0f / det;
  f3 tvec = o - p1;
  float uu = dot3(tvec, pvec) * inv_det;
  if (uu < 0.0f || uu > 1.0f) return false;
  f3 qvec = cross3(tvec, edge1);
  float vv = dot3(d, qvec) * inv_det;
  if (vv < 0.0f || uu + vv > 1.0f) return false;
  float tt = dot3(edge2, qvec) * inv_det;
  if (tt < tmin || tt > tmax) return false;
  *t = tt;
  *u = uu;
  *v = vv;
  return true;
}

// ===============================================================================================
// MODE_REF
// ===============================================================================================
#define JT_REF_STACK 128 /* the reference's default --bvhstacksize (src/cli.jl:82-85) */

// intersect_bbox, src/geometry.jl:96-105
JT_DEV bool ref_bbox(f3 o, f3 dinv, float tmin, float tmax, f3 bmin, f3 bmax) {
  f3 it_min = (bmin - o) * dinv;
  f3 it_max = (bmax - o) * dinv;
  float lo = jl_max(jl_max(jl_min(it_min.x, it_max.x), jl_min(it_min.y, it_max.y)), jl_min(it_min.z, it_max.z));
  float hi = jl_min(jl_min(jl_max(it_min.x, it_max.x), jl_max(it_min.y, it_max.y)), jl_max(it_min.z, it_max.z));
  float t0 = jl_max(lo, tmin);
  float t1 = jl_min(hi, tmax);
  double t1d = (double)t1 * 1.00000024;  // Q3
  return (double)t0 <= t1d;
}

// intersect_shape_bvh on one shape, ray already in instance space. Updates best (t,u,v,elem).
JT_DEV bool ref_shape(const JtDevScene& S, const JtShapeRec& sh, f3 o, f3 d, float tmin, float tmax, float* bt,
                      float* bu, float* bv, int* belem) {
  if (sh.num_ref_nodes == 0) return false;
  int stack[JT_REF_STACK];
  int sp = 0;
  stack[sp++] = 0;
  bool hit = false;
  f3 dinv = f3{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
  const float4* nodes = S.ref_nodes + 2 * (size_t)sh.ref_node_off;
  const int32_t* prims = S.ref_prims + sh.ref_prim_off;
  const float* pos = S.positions + 3 * (size_t)sh.pos_off;
  const int4* elems = S.elements + sh.elem_off;
  while (sp != 0) {
    int ni = stack[--sp];
    float4 a = __ldg(nodes + 2 * ni), b = __ldg(nodes + 2 * ni + 1);
    if (!ref_bbox(o, dinv, tmin, tmax, f3{a.x, a.y, a.z}, f3{a.w, b.x, b.y})) continue;
    int start = __float_as_int(b.z);
    int packed = __float_as_int(b.w);
    int num = packed & 0xFFFF, axis = (packed >> 16) & 0xFF, internal = packed >> 24;
    if (internal) {
      float da = axis == 0 ? d.x : (axis == 1 ? d.y : d.z);
      if (!(da < 0.0f)) {  // ray_dsign == 0: push start, start+1 -> start+1 popped first (Q1)
        stack[sp++] = start;
        stack[sp++] = start + 1;
      } else {
        stack[sp++] = start + 1;
        stack[sp++] = start;
      }
    } else {
      for (int i = start; i < start + num; i++) {
        int e = __ldg(prims + i);
        int4 q = __ldg(elems + e);
        f3 p1 = ld3(pos + 3 * q.x), p2 = ld3(pos + 3 * q.y);
        float t, u, v;
        if (sh.kind == 1) {
          f3 p3 = ld3(pos + 3 * q.z);
          if (!tri_test(o, d, tmin, tmax, p1, p2 - p1, p3 - p1, &t, &u, &v)) continue;
        } else {
          // intersect_quad, src/geometry.jl:238-258
          f3 p3 = ld3(pos + 3 * q.z), p4 = ld3(pos + 3 * q.w);
          float t1 = 0, u1 = 0, v1 = 0, t2 = 0, u2 = 0, v2 = 0;
          bool h1 = tri_test(o, d, tmin, tmax, p1, p2 - p1, p4 - p1, &t1, &u1, &v1);
          if (p3 == p4) {
            if (!h1) continue;
            t = t1; u = u1; v = v1;
          } else {
            bool h2 = tri_test(o, d, tmin, tmax, p3, p4 - p3, p2 - p3, &t2, &u2, &v2);
            // isec1.distance < isec2.distance ? isec1 : isec2, misses carry distance = +Inf
            float d1 = h1 ? t1 : INFINITY, d2 = h2 ? t2 : INFINITY;
            if (d1 < d2) {
              t = t1; u = u1; v = v1;
            } else {
              if (!h2) continue;
              t = t2; u = 1.0f - u2; v = 1.0f - v2;
            }
          }
        }
        *bt = t; *bu = u; *bv = v; *belem = e;  // Q2: every accepted hit overwrites
        tmax = t;
        hit = true;
      }
    }
  }
  return hit;
}

// intersect_instance_bvh
JT_DEV DHit ref_instance(const JtDevScene& S, int inst, const DRay& ray) {
  const JtInstanceRec& I = S.instances[inst];
  f3 o = xform_point(I.inv, ray.o), d = xform_vector(I.inv, ray.d);
  DHit h{0.0f, 0.0f, 0.0f, -1, -1};
  float t, u, v;
  int e;
  if (ref_shape(S, S.shapes[I.shape], o, d, ray.tmin, ray.tmax, &t, &u, &v, &e)) h = DHit{t, u, v, inst, e};
  return h;
}

// intersect_scene_bvh
JT_DEV DHit ref_scene(const JtDevScene& S, const DRay& ray) {
  DHit best{0.0f, 0.0f, 0.0f, -1, -1};
  if (S.tlas_num_nodes == 0) return best;
  int stack[JT_REF_STACK];
  int sp = 0;
  stack[sp++] = 0;
  f3 o = ray.o, d = ray.d;
  float tmax = ray.tmax;
  f3 dinv = f3{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
  while (sp != 0) {
    int ni = stack[--sp];
    float4 a = __ldg(S.ref_nodes + 2 * ni), b = __ldg(S.ref_nodes + 2 * ni + 1);
    if (!ref_bbox(o, dinv, ray.tmin, tmax, f3{a.x, a.y, a.z}, f3{a.w, b.x, b.y})) continue;
    int start = __float_as_int(b.z);
    int packed = __float_as_int(b.w);
    int num = packed & 0xFFFF, axis = (packed >> 16) & 0xFF, internal = packed >> 24;
    if (internal) {
      float da = axis == 0 ? d.x : (axis == 1 ? d.y : d.z);
      if (!(da < 0.0f)) {
        stack[sp++] = start;
        stack[sp++] = start + 1;
      } else {
        stack[sp++] = start + 1;
        stack[sp++] = start;
      }
    } else {
      for (int i = start; i < start + num; i++) {
        int inst = __ldg(S.ref_prims + i);
        const JtInstanceRec& I = S.instances[inst];
        f3 io = xform_point(I.inv, o), id = xform_vector(I.inv, d);
        float t, u, v;
        int e;
        if (!ref_shape(S, S.shapes[I.shape], io, id, ray.tmin, tmax, &t, &u, &v, &e)) continue;
        best = DHit{t, u, v, inst, e};
        tmax = t;
      }
    }
  }
  return best;
}

// ===============================================================================================
// MODE_WIDE
// ===============================================================================================
#define JT_WIDE_STACK 64 /* entries; staging rejects scenes whose wide BVH could need more */

struct WideRay {  // per-space traversal constants
  f3 o, d;
  float idx, idy, idz;  // guarded reciprocal direction for the slab tests
  uint32_t oct;         // slot index pointing along the ray: bit k set <=> d[k] >= 0
  uint32_t rank_oct;    // reference octant: bit k set <=> d[k] < 0
};

JT_DEV float guarded_rcp(float d) {
  const float tiny = 1e-30f;
  float a = fabsf(d) < tiny ? copysignf(tiny, d) : d;
  return 1.0f / a;
}
JT_DEV void wide_ray_setup(WideRay& R, f3 o, f3 d) {
  R.o = o;
  R.d = d;
  R.idx = guarded_rcp(d.x);
  R.idy = guarded_rcp(d.y);
  R.idz = guarded_rcp(d.z);
  uint32_t neg = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
  R.rank_oct = neg;
  R.oct = neg ^ 7u;
}

JT_DEV uint32_t byte_of(uint32_t w, int i) { return (w >> (8 * i)) & 0xFFu; }
// Byte i of w as the float 256 + b/128 without an I2F (1/8 of the FP32 rate; the node test needs 48
// conversions): one PRMT splices the byte into mantissa bits 8..15 of 256.0f. The affine map back to b is
// folded into the slab FMA: t = b*s + a = (256 + b/128)*(128 s) + (a - 32768 s).
// SASS PRMT takes ONE non-register operand. With the literal 0x43800000 ptxas spends it on that and re-materialises the
// four selectors into registers before every use (48 extra IMAD.U32 / MOV per node step in the round-1 build). Read from
// constant memory the pattern is opaque, lives in one register, and the selector becomes the immediate.
#if defined(__CUDACC__) && !defined(JT_EMU_COUNT)
__constant__ uint32_t jt_c_prmt_magic = 0x43800000u;
#endif
#if defined(__CUDA_ARCH__) && !defined(JT_EMU_COUNT)
#define JT_PRMT_MAGIC jt_c_prmt_magic
#else
#define JT_PRMT_MAGIC 0x43800000u
#endif
// byte k of the result = 0xFF if bit 31 of word k is set, else 0x00
JT_DEV uint32_t sign_bytes4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
#if defined(__CUDA_ARCH__) && !defined(JT_EMU_COUNT)
  uint32_t ab, cd, r;
  // selector nibble = 8 | byte index: replicate that byte's most significant bit
  asm("prmt.b32 %0, %1, %2, 0x00FB;" : "=r"(ab) : "r"(a), "r"(b));
  asm("prmt.b32 %0, %1, %2, 0x00FB;" : "=r"(cd) : "r"(c), "r"(d));
  asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(r) : "r"(ab), "r"(cd));
  return r;
#else
  return ((a >> 31) * 0xFFu) | ((b >> 31) * 0xFF00u) | ((c >> 31) * 0xFF0000u) | ((d >> 31) * 0xFF000000u);
#endif
}
JT_DEV float byte_as_biased_float(uint32_t w, int i, uint32_t magic) {
  return __uint_as_float(__byte_perm(w, magic, 0x7604u | ((uint32_t)i << 4)));
}

// Slab-test the 8 children of one node. Returns the 32-bit hit word: bits 24..31 internal
// children in priority order, bits 0..23 triangle records.
JT_DEV uint32_t wide_node_hits(const float4* __restrict__ wnodes, uint32_t node, const WideRay& R, float tmin,
                               float tmax, uint2* ngroup_out, uint2* tgroup_out) {
  JT_COUNT(wide_nodes);
  const float4* np = wnodes + 5 * (size_t)node;
  float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
  uint32_t eimask = __float_as_uint(n0.w);
  // grid step scaled by the reciprocal direction; origin term; absolute slack covering the
  // rounding of this (fused) formulation so the test never rejects what exact arithmetic accepts
  float sx = __uint_as_float((eimask & 0xFFu) << 23) * R.idx;
  float sy = __uint_as_float(((eimask >> 8) & 0xFFu) << 23) * R.idy;
  float sz = __uint_as_float(((eimask >> 16) & 0xFFu) << 23) * R.idz;
  float ax = (n0.x - R.o.x) * R.idx, ay = (n0.y - R.o.y) * R.idy, az = (n0.z - R.o.z) * R.idz;
  // slack: 2^-21 of the magnitudes involved + 2^-7 of a grid step for the biased-byte folding below
  const float k = 4.76837158e-7f;  // 2^-21
  float ex = fmaf(k, fmaf(256.0f, fabsf(sx), fabsf(ax)), 0.0078125f * fabsf(sx));
  float ey = fmaf(k, fmaf(256.0f, fabsf(sy), fabsf(ay)), 0.0078125f * fabsf(sy));
  float ez = fmaf(k, fmaf(256.0f, fabsf(sz), fabsf(az)), 0.0078125f * fabsf(sz));
  // fold the byte bias: b = 128 f - 32768
  float alx = fmaf(-32768.0f, sx, ax) - ex, ahx = fmaf(-32768.0f, sx, ax) + ex;
  float aly = fmaf(-32768.0f, sy, ay) - ey, ahy = fmaf(-32768.0f, sy, ay) + ey;
  float alz = fmaf(-32768.0f, sz, az) - ez, ahz = fmaf(-32768.0f, sz, az) + ez;
  sx *= 128.0f;
  sy *= 128.0f;
  sz *= 128.0f;
  uint32_t imask = eimask >> 24;
  uint32_t meta_w[2] = {__float_as_uint(n1.z), __float_as_uint(n1.w)};
  uint32_t qlox[2] = {__float_as_uint(n2.x), __float_as_uint(n2.y)};
  uint32_t qloy[2] = {__float_as_uint(n2.z), __float_as_uint(n2.w)};
  uint32_t qloz[2] = {__float_as_uint(n3.x), __float_as_uint(n3.y)};
  uint32_t qhix[2] = {__float_as_uint(n3.z), __float_as_uint(n3.w)};
  uint32_t qhiy[2] = {__float_as_uint(n4.x), __float_as_uint(n4.y)};
  uint32_t qhiz[2] = {__float_as_uint(n4.z), __float_as_uint(n4.w)};
  const bool px = R.d.x >= 0.0f, py = R.d.y >= 0.0f, pz = R.d.z >= 0.0f;
  const uint32_t oct4 = R.oct * 0x01010101u;
  const uint32_t magic = JT_PRMT_MAGIC;
  uint32_t hits = 0;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    uint32_t nearx = px ? qlox[h] : qhix[h], farx = px ? qhix[h] : qlox[h];
    uint32_t neary = py ? qloy[h] : qhiy[h], fary = py ? qhiy[h] : qloy[h];
    uint32_t nearz = pz ? qloz[h] : qhiz[h], farz = pz ? qhiz[h] : qloz[h];
    // decode the four meta bytes at once (branch-free, after Ylitie et al. 2017):
    //   internal child: 0b001_11sss -> bit 24 + (slot ^ octant), 1 bit;  leaf: count<<5 | offset -> bits at offset;
    //   empty slot: 0 -> contributes no bits
    const uint32_t meta4 = meta_w[h];
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = (is_inner4 >> 4) * 7u;  // 0x07 in every internal byte
    const uint32_t bit_index4 = (meta4 ^ (oct4 & inner_mask4)) & 0x1F1F1F1Fu;
    // The ray misses child j iff  max(t0x, t0y, t0z, tmin) > min(t1x, t1y, t1z, tmax), i.e. iff one of the three
    // differences below is negative (none is ever NaN: the slab values are finite, tmin is, tmax is finite or +inf).
    // Taking the clamps out of the min / max chains moves two operations per child from the ALU pipe (FMNMX), which
    // bounds this loop together with the PRMT conversions, to the FMA pipe (FADD); the sign bits are OR-ed by one LOP3.
    uint32_t sgn[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float t0x = fmaf(byte_as_biased_float(nearx, j, magic), sx, alx), t1x = fmaf(byte_as_biased_float(farx, j, magic), sx, ahx);
      float t0y = fmaf(byte_as_biased_float(neary, j, magic), sy, aly), t1y = fmaf(byte_as_biased_float(fary, j, magic), sy, ahy);
      float t0z = fmaf(byte_as_biased_float(nearz, j, magic), sz, alz), t1z = fmaf(byte_as_biased_float(farz, j, magic), sz, ahz);
      float lo = fmaxf(fmaxf(t0x, t0y), t0z);
      float hi = fminf(fminf(t1x, t1y), t1z);
      sgn[j] = __float_as_uint(hi - lo) | __float_as_uint(tmax - lo) | __float_as_uint(hi - tmin);
    }
    // the four sign bits, smeared over their bytes by PRMT's sign-replication mode, knock the missed children's bit
    // counts out of the packed word at once (3 PRMT + 1 LOP3 instead of 4 FSETP + 4 SEL)
    const uint32_t miss4 = sign_bytes4(sgn[0], sgn[1], sgn[2], sgn[3]);
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u & ~miss4;
#pragma unroll
    for (int j = 0; j < 4; j++) hits |= __byte_perm(child_bits4, 0u, 0x4440u + (uint32_t)j) << (byte_of(bit_index4, j) & 31u);
  }
  *ngroup_out = make_uint2(__float_as_uint(n1.x), (hits & 0xFF000000u) | imask);
  *tgroup_out = make_uint2(__float_as_uint(n1.y), hits & 0x00FFFFFFu);
  return hits;
}

struct WideBest {
  float t, u, v;
  int inst, elem;
  int wtri;  // record index of the winning triangle (for tie ranks)
};

// Candidate (t,u,v) from record `wtri` of instance `inst`: closest wins; an exact-t tie goes to
// the candidate the reference would have visited LAST (Q1 + Q2) -- looked up in the rank tables.
JT_DEV void wide_accept(const JtDevScene& S, WideBest& B, float t, float u, float v, int inst, int elem, int wtri,
                        uint32_t world_oct, uint32_t local_oct) {
  if (B.inst >= 0 && t == B.t) {
    bool later;
    if (inst != B.inst) {
      later = __ldg(S.inst_rank + (size_t)world_oct * S.num_instances + inst) >
              __ldg(S.inst_rank + (size_t)world_oct * S.num_instances + B.inst);
    } else {
      later = __ldg(S.tri_rank + (size_t)local_oct * S.num_wtris + wtri) >
              __ldg(S.tri_rank + (size_t)local_oct * S.num_wtris + B.wtri);
    }
    if (!later) return;
  }
  B.t = t; B.u = u; B.v = v; B.inst = inst; B.elem = elem; B.wtri = wtri;
}

// Leaf record of a FLATTENED instance (flags bit 9): its triangle is stored in instance space, so the world ray is
// taken into that space exactly as intersect_scene_bvh does per instance visit (transform_ray with the precomputed
// inverse(frame, true), src/bvh.jl:345-348) and the reference's own leaf arithmetic runs unchanged.
JT_DEV void flat_ray(const JtDevScene& S, int inst, f3 wo, f3 wd, f3& o, f3& d, uint32_t& local_oct) {
  JT_COUNT(wide_xforms);
  const JtInstanceRec& I = S.instances[inst];
  o = xform_point(I.inv, wo);
  d = xform_vector(I.inv, wd);
  local_oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
}

// Shared traversal core. `root` is a wide node index; when `single_inst` >= 0 the walk starts
// inside that instance's BLAS with the ray already in instance space (intersect_instance_bvh).
JT_DEV void wide_walk(const JtDevScene& S, uint32_t root, f3 o, f3 d, float tmin, float tmax, int single_inst,
                      WideBest& B) {
  uint2 stack[JT_WIDE_STACK];
  int sp = 0;
  WideRay R, Rworld;
  wide_ray_setup(R, o, d);
  Rworld = R;
  const uint32_t world_oct = R.rank_oct;
  int cur_inst = single_inst;  // >= 0 while inside an instanced BLAS
  int blas_sp = -1;            // stack height at BLAS entry (-1 = in the top level / single-instance walk)
  uint2 ngroup = make_uint2(root, 0x80000000u), tgroup = make_uint2(0u, 0u);
  float best_t = tmax;  // current acceptance bound (reference: ray.tmax shrinks with every hit)
  for (;;) {
    if (ngroup.y > 0x00FFFFFFu) {
      uint32_t hits = ngroup.y;
      uint32_t bit = 31u - (uint32_t)__clz(hits);
      hits &= ~(1u << bit);
      ngroup.y = hits;
      if (hits > 0x00FFFFFFu) stack[sp++] = ngroup;
      uint32_t slot = (bit - 24u) ^ R.oct;
      uint32_t rel = __popc(hits & 0xFFu & ~(0xFFFFFFFFu << slot));
      wide_node_hits(S.wnodes, ngroup.x + rel, R, tmin, best_t, &ngroup, &tgroup);
    } else {
      tgroup = ngroup;
      ngroup = make_uint2(0u, 0u);
    }
    while (tgroup.y != 0u) {
      uint32_t bit = (uint32_t)__ffs((int)tgroup.y) - 1u;
      tgroup.y &= ~(1u << bit);
      uint32_t wtri = tgroup.x + bit;
      const float4* tp = S.wtris + 3 * (size_t)wtri;
      float4 r0 = __ldg(tp), r1 = __ldg(tp + 1), r2 = __ldg(tp + 2);
      uint32_t flags = __float_as_uint(r2.w);
      JT_COUNT(wide_prims);
      if (flags & 0x100u) {
        JT_COUNT(wide_instances);
        // instance record: park the rest of the world-level work, switch to instance space
        int inst = __float_as_int(r1.w);
        if (tgroup.y != 0u) stack[sp++] = tgroup;
        if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;
        blas_sp = sp;
        cur_inst = inst;
        const JtInstanceRec& I = S.instances[inst];
        wide_ray_setup(R, xform_point(I.inv, Rworld.o), xform_vector(I.inv, Rworld.d));
        const int entry = __float_as_int(r0.w);  // braided sub-tree of the BLAS, or -1: the whole shape
        ngroup = make_uint2((uint32_t)(entry >= 0 ? entry : S.shapes[I.shape].wide_root), 0x80000000u);
        tgroup = make_uint2(0u, 0u);
        break;
      }
      float t, u, v;
      f3 to = R.o, td = R.d;
      uint32_t local_oct = R.rank_oct;
      if (flags & 0x200u) flat_ray(S, __float_as_int(r1.w), Rworld.o, Rworld.d, to, td, local_oct);
      if (tri_test(to, td, tmin, best_t, f3{r0.x, r0.y, r0.z}, f3{r1.x, r1.y, r1.z}, f3{r2.x, r2.y, r2.z}, &t,
                   &u, &v)) {
        int inst = cur_inst >= 0 ? cur_inst : __float_as_int(r1.w);
        if (flags & 1u) {
          u = 1.0f - u;
          v = 1.0f - v;
        }
        wide_accept(S, B, t, u, v, inst, __float_as_int(r0.w), (int)wtri, world_oct, local_oct);
        best_t = B.t;
      }
    }
    if (ngroup.y <= 0x00FFFFFFu) {
      if (sp == blas_sp) {  // the BLAS is exhausted: back to world space
        blas_sp = -1;
        cur_inst = -1;
        R = Rworld;
      }
      if (sp == 0) break;
      ngroup = stack[--sp];
    }
  }
}

JT_DEV DHit wide_scene(const JtDevScene& S, const DRay& ray) {
  WideBest B{0.0f, 0.0f, 0.0f, -1, -1, -1};
  if (S.wide_root >= 0) wide_walk(S, (uint32_t)S.wide_root, ray.o, ray.d, ray.tmin, ray.tmax, -1, B);
  return DHit{B.t, B.u, B.v, B.inst, B.elem};
}

JT_DEV DHit wide_instance(const JtDevScene& S, int inst, const DRay& ray) {
  const JtInstanceRec& I = S.instances[inst];
  WideBest B{0.0f, 0.0f, 0.0f, -1, -1, -1};
  int root = S.shapes[I.shape].wide_root;
  if (root >= 0)
    wide_walk(S, (uint32_t)root, xform_point(I.inv, ray.o), xform_vector(I.inv, ray.d), ray.tmin, ray.tmax, inst, B);
  return DHit{B.t, B.u, B.v, B.inst, B.elem};
}

// ---- mode dispatch -------------------------------------------------------------------------------
template <int MODE>
JT_DEV DHit intersect_scene(const JtDevScene& S, const DRay& ray) {
  if (MODE == MODE_REF) return ref_scene(S, ray);
  return wide_scene(S, ray);
}
template <int MODE>
JT_DEV DHit intersect_instance(const JtDevScene& S, int inst, const DRay& ray) {
  if (MODE == MODE_REF) return ref_instance(S, inst, ray);
  return wide_instance(S, inst, ray);
}
