// jt_host_scene.cpp -- the host steps around the hot path, natively (SURVEY.md 8f N2 + the host half of N1 / N4):
//   load_scene        src/sceneio.jl:25-93     JSON scene description, defaults of src/scene.jl:58-263
//   load_shape        src/shape.jl:78-124, :302-446   PLY (binary little endian / ascii), quad promotion, fan
//                                              triangulation, v-flip of texture coordinates
//   load_texture      src/scene.jl:164-189     8-bit PNG (own decoder over zlib's inflate), Radiance RGBE .hdr
//   make_scene_bvh    src/bvh.jl:66-136        primitive / instance boxes + jt_make_bvh (jt_host_bvh.cpp)
//   make_trace_lights src/trace.jl:117-187     sequential Float32 CDFs (area lights, textured environments)
// A jt_host_scene owns every array and exposes the jt_scene_desc that jt_scene_create / jt_group_create consume, so a
// C or C++ host needs nothing but this library to go from a scenes/<name>/<name>.json to a rendered image.
// The Python mirror (sceneio.py, bvh.py, lights.py) produces bit-identical arrays: tests/test_native_host.py.
//
// Missing-asset rule: an absent texture file becomes a 1x1 opaque white RGBA8 texture, instances of an
// absent shape file are dropped before BVH and light building; every substitution is recorded as a note.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <zlib.h>

#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "jt_internal.h"

namespace {

struct Fail : std::runtime_error {
  int code;
  Fail(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] void fail(int code, const std::string& msg) { throw Fail(code, msg); }

bool file_exists(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

std::vector<uint8_t> read_file(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) fail(JT_ERR_INVALID, "cannot open " + path);
  std::vector<uint8_t> buf;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize(n > 0 ? (size_t)n : 0);
  if (n > 0 && fread(buf.data(), 1, (size_t)n, f) != (size_t)n) {
    fclose(f);
    fail(JT_ERR_INVALID, "short read on " + path);
  }
  fclose(f);
  return buf;
}

// ---------------------------------------------------------------------------------------------------------------------
// JSON (RFC 8259 subset: everything the scene files use; numbers through strtod like Python's float())
// ---------------------------------------------------------------------------------------------------------------------
struct JVal {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<JVal> arr;
  std::vector<std::pair<std::string, JVal>> obj;
  const JVal* get(const char* key) const {
    if (kind != Obj) return nullptr;
    for (const auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
};

struct JParser {
  const char* p;
  const char* end;
  const std::string& where;
  void ws() {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
  }
  [[noreturn]] void bad(const char* what) { fail(JT_ERR_INVALID, where + ": JSON " + what); }
  JVal value() {
    ws();
    if (p >= end) bad("ends early");
    JVal v;
    if (*p == '{') {
      v.kind = JVal::Obj;
      p++;
      ws();
      if (p < end && *p == '}') { p++; return v; }
      for (;;) {
        ws();
        if (p >= end || *p != '"') bad("object key expected");
        std::string k = string();
        ws();
        if (p >= end || *p != ':') bad("':' expected");
        p++;
        v.obj.emplace_back(std::move(k), value());
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == '}') { p++; return v; }
        bad("',' or '}' expected");
      }
    }
    if (*p == '[') {
      v.kind = JVal::Arr;
      p++;
      ws();
      if (p < end && *p == ']') { p++; return v; }
      for (;;) {
        v.arr.push_back(value());
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == ']') { p++; return v; }
        bad("',' or ']' expected");
      }
    }
    if (*p == '"') {
      v.kind = JVal::Str;
      v.str = string();
      return v;
    }
    if (end - p >= 4 && !strncmp(p, "true", 4)) { v.kind = JVal::Bool; v.b = true; p += 4; return v; }
    if (end - p >= 5 && !strncmp(p, "false", 5)) { v.kind = JVal::Bool; v.b = false; p += 5; return v; }
    if (end - p >= 4 && !strncmp(p, "null", 4)) { p += 4; return v; }
    char* q = nullptr;
    v.num = strtod(p, &q);
    if (q == p) bad("value expected");
    v.kind = JVal::Num;
    p = q;
    return v;
  }
  std::string string() {
    std::string s;
    p++;  // opening quote
    while (p < end && *p != '"') {
      if (*p == '\\') {
        p++;
        if (p >= end) bad("bad escape");
        switch (*p) {
          case 'n': s += '\n'; break;
          case 't': s += '\t'; break;
          case 'r': s += '\r'; break;
          case 'b': s += '\b'; break;
          case 'f': s += '\f'; break;
          case 'u': {
            if (end - p < 5) bad("bad \\u escape");
            unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
            p += 4;
            if (cp < 0x80) s += (char)cp;
            else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
            else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: s += *p;  // \" \\ \/
        }
        p++;
      } else {
        s += *p++;
      }
    }
    if (p >= end) bad("unterminated string");
    p++;
    return s;
  }
};

float jnum(const JVal* o, const char* key, double dflt) {
  const JVal* v = o->get(key);
  return (float)((v && v->kind == JVal::Num) ? v->num : dflt);  // np.float32(python float): round to nearest
}
void jvec3(const JVal* o, const char* key, float out[3]) {
  const JVal* v = o->get(key);
  out[0] = out[1] = out[2] = 0.0f;
  if (v && v->kind == JVal::Arr)
    for (size_t k = 0; k < 3 && k < v->arr.size(); k++) out[k] = (float)v->arr[k].num;
}
int64_t jid(const JVal* o, const char* key) {  // 0-based in the file -> 1-based, missing -> -1 (src/scene.jl:45)
  const JVal* v = o->get(key);
  return (v && v->kind == JVal::Num) ? (int64_t)v->num + 1 : -1;
}
// Frame3f(array), src/math.jl:47-60: 12 floats -> x, y, z, o columns; any other length -> identity
void jframe(const JVal* o, jt_frame* f) {
  static const float ident[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
  const JVal* v = o->get("frame");
  float* dst = f->x;
  if (v && v->kind == JVal::Arr && v->arr.size() == 12)
    for (int k = 0; k < 12; k++) dst[k] = (float)v->arr[(size_t)k].num;
  else
    memcpy(dst, ident, sizeof(ident));
}
void no_lookat(const JVal* o, const char* what) {
  if (o->get("lookat")) fail(JT_ERR_UNSUPPORTED, std::string(what) + " 'lookat' is not used by any shipped scene; unsupported");
}

// ---------------------------------------------------------------------------------------------------------------------
// shapes: PLY, src/shape.jl:78-124
// ---------------------------------------------------------------------------------------------------------------------
struct HostShape {
  std::vector<float> positions, normals, texcoords, colors;  // 3 / 3 / 2 / 4 per vertex
  std::vector<int64_t> triangles, quads;                     // 1-based
  std::vector<jt_bvh_node> nodes;
  std::vector<int64_t> prims;
  bool empty() const { return triangles.empty() && quads.empty(); }
};

struct PlyProp {
  bool is_list = false;
  std::string name;
  int type = 0, count_type = 0;  // index into kPlyTypes
};
struct PlyElement {
  std::string name;
  int64_t count = 0;
  std::vector<PlyProp> props;
};
struct PlyType { const char* a; const char* b; int size; char kind; };  // kind: i, u, f
const PlyType kPlyTypes[] = {{"char", "int8", 1, 'i'},   {"uchar", "uint8", 1, 'u'},  {"short", "int16", 2, 'i'},
                             {"ushort", "uint16", 2, 'u'}, {"int", "int32", 4, 'i'},    {"uint", "uint32", 4, 'u'},
                             {"float", "float32", 4, 'f'}, {"double", "float64", 8, 'f'}};
int ply_type(const std::string& t, const std::string& path) {
  for (int k = 0; k < 8; k++)
    if (t == kPlyTypes[k].a || t == kPlyTypes[k].b) return k;
  fail(JT_ERR_INVALID, path + ": unknown PLY type " + t);
}
double ply_read(const uint8_t* p, int type) {
  switch (type) {
    case 0: { int8_t v; memcpy(&v, p, 1); return v; }
    case 1: return *p;
    case 2: { int16_t v; memcpy(&v, p, 2); return v; }
    case 3: { uint16_t v; memcpy(&v, p, 2); return v; }
    case 4: { int32_t v; memcpy(&v, p, 4); return v; }
    case 5: { uint32_t v; memcpy(&v, p, 4); return v; }
    case 6: { float v; memcpy(&v, p, 4); return v; }
    default: { double v; memcpy(&v, p, 8); return v; }
  }
}

// `get_faces` (src/shape.jl:302-446): if any face has exactly 4 indices the whole shape is stored as quads (a triangle
// becomes (a, b, c, c), an n-gon a fan of degenerate quads), else as fan-triangulated triangles. 0-based in, 1-based out.
void faces_to_elements(const std::vector<int64_t>& vals, const std::vector<int64_t>& starts, HostShape* sh) {
  const size_t nf = starts.size() - 1;
  bool has_quads = false;
  for (size_t i = 0; i < nf; i++) has_quads = has_quads || (starts[i + 1] - starts[i] == 4);
  const int width = has_quads ? 4 : 3;
  std::vector<int64_t>& out = has_quads ? sh->quads : sh->triangles;
  for (size_t i = 0; i < nf; i++) {
    const int64_t* d = vals.data() + starts[i];
    const int64_t n = starts[i + 1] - starts[i];
    if (n < 3) {
      for (int k = 0; k < width; k++) out.push_back((k < n ? d[k] : -1) + 1);
    } else if (n == 3) {
      out.push_back(d[0] + 1); out.push_back(d[1] + 1); out.push_back(d[2] + 1);
      if (has_quads) out.push_back(d[2] + 1);
    } else if (n == 4 && has_quads) {
      for (int k = 0; k < 4; k++) out.push_back(d[k] + 1);
    } else {
      for (int64_t item = 2; item < n; item++) {
        out.push_back(d[0] + 1); out.push_back(d[item - 1] + 1); out.push_back(d[item] + 1);
        if (has_quads) out.push_back(d[item] + 1);
      }
    }
  }
}

void load_shape(const std::string& path, HostShape* sh) {
  std::vector<uint8_t> data = read_file(path);
  const char* text = (const char*)data.data();
  const char* eh = (const char*)memmem(data.data(), data.size(), "end_header", 10);
  if (!eh || data.size() < 4 || memcmp(text, "ply", 3) != 0) fail(JT_ERR_INVALID, path + ": not a PLY file");
  size_t body = (size_t)(eh - text) + 10;
  if (body + 1 < data.size() && text[body] == '\r' && text[body + 1] == '\n') body += 2;
  else body += 1;
  std::string format;
  std::vector<PlyElement> elements;
  {
    std::string header(text, (size_t)(eh - text));
    size_t pos = 0;
    while (pos < header.size()) {
      size_t nl = header.find('\n', pos);
      if (nl == std::string::npos) nl = header.size();
      std::string line = header.substr(pos, nl - pos);
      pos = nl + 1;
      std::vector<std::string> tok;
      size_t i = 0;
      while (i < line.size()) {
        while (i < line.size() && isspace((unsigned char)line[i])) i++;
        size_t j = i;
        while (j < line.size() && !isspace((unsigned char)line[j])) j++;
        if (j > i) tok.push_back(line.substr(i, j - i));
        i = j;
      }
      if (tok.empty()) continue;
      if (tok[0] == "format" && tok.size() >= 2) {
        format = tok[1];
      } else if (tok[0] == "element" && tok.size() >= 3) {
        PlyElement e;
        e.name = tok[1];
        e.count = atoll(tok[2].c_str());
        elements.push_back(e);
      } else if (tok[0] == "property" && !elements.empty()) {
        PlyProp p;
        if (tok.size() >= 5 && tok[1] == "list") {
          p.is_list = true;
          p.count_type = ply_type(tok[2], path);
          p.type = ply_type(tok[3], path);
          p.name = tok[4];
        } else if (tok.size() >= 3) {
          p.type = ply_type(tok[1], path);
          p.name = tok[2];
        } else {
          fail(JT_ERR_INVALID, path + ": bad property line");
        }
        elements.back().props.push_back(p);
      }
    }
  }
  const bool binary = format == "binary_little_endian";
  if (!binary && format != "ascii") fail(JT_ERR_UNSUPPORTED, path + ": unsupported PLY format " + format);

  std::map<std::string, std::vector<float>> vert;  // scalar vertex properties, converted like np.asarray(.., float32)
  std::vector<std::string> vnames;
  std::vector<int64_t> face_vals, face_starts;
  bool have_faces = false;

  size_t pos = body;                               // binary cursor
  const char* ap = text + body;                    // ascii cursor
  const char* aend = text + data.size();
  auto ascii_token = [&]() -> double {
    while (ap < aend && isspace((unsigned char)*ap)) ap++;
    if (ap >= aend) fail(JT_ERR_INVALID, path + ": PLY body ends early");
    char* q = nullptr;
    double v = strtod(ap, &q);
    if (q == ap) fail(JT_ERR_INVALID, path + ": bad number in PLY body");
    ap = q;
    return v;
  };
  for (const PlyElement& el : elements) {
    bool all_scalar = true;
    for (const PlyProp& p : el.props) all_scalar = all_scalar && !p.is_list;
    const bool single_list = el.props.size() == 1 && el.props[0].is_list;
    if (!all_scalar && !single_list) fail(JT_ERR_UNSUPPORTED, path + ": mixed scalar/list element '" + el.name + "' not supported");
    if ((el.name == "line" || el.name == "point") && el.count > 0)
      fail(JT_ERR_UNSUPPORTED, path + ": '" + el.name + "' elements crash the reference (SURVEY.md 2.3); not supported");
    const bool is_vertex = el.name == "vertex", is_face = el.name == "face";
    if (all_scalar) {
      std::vector<std::vector<float>*> cols;
      if (is_vertex)
        for (const PlyProp& p : el.props) {
          vnames.push_back(p.name);
          vert[p.name].assign((size_t)el.count, 0.0f);
          cols.push_back(&vert[p.name]);
        }
      size_t rec = 0;
      for (const PlyProp& p : el.props) rec += (size_t)kPlyTypes[p.type].size;
      if (binary) {
        if (pos + rec * (size_t)el.count > data.size()) fail(JT_ERR_INVALID, path + ": PLY body ends early");
        if (is_vertex)
          for (int64_t i = 0; i < el.count; i++) {
            const uint8_t* r = data.data() + pos + rec * (size_t)i;
            for (size_t k = 0; k < el.props.size(); k++) {
              (*cols[k])[(size_t)i] = (float)ply_read(r, el.props[k].type);
              r += kPlyTypes[el.props[k].type].size;
            }
          }
        pos += rec * (size_t)el.count;
      } else {
        for (int64_t i = 0; i < el.count; i++)
          for (size_t k = 0; k < el.props.size(); k++) {
            double v = ascii_token();
            if (is_vertex) (*cols[k])[(size_t)i] = (float)v;
          }
      }
    } else {
      const PlyProp& p = el.props[0];
      const bool keep = is_face && p.name == "vertex_indices";
      if (keep) {
        have_faces = true;
        face_starts.push_back(0);
      }
      const int csz = kPlyTypes[p.count_type].size, isz = kPlyTypes[p.type].size;
      for (int64_t i = 0; i < el.count; i++) {
        int64_t n;
        if (binary) {
          if (pos + (size_t)csz > data.size()) fail(JT_ERR_INVALID, path + ": PLY body ends early");
          n = (int64_t)ply_read(data.data() + pos, p.count_type);
          pos += (size_t)csz;
          if (n < 0 || pos + (size_t)(n * isz) > data.size()) fail(JT_ERR_INVALID, path + ": PLY body ends early");
          for (int64_t k = 0; k < n; k++) {
            if (keep) face_vals.push_back((int64_t)ply_read(data.data() + pos, p.type));
            pos += (size_t)isz;
          }
        } else {
          n = (int64_t)ascii_token();
          for (int64_t k = 0; k < n; k++) {
            double v = ascii_token();
            if (keep) face_vals.push_back((int64_t)v);
          }
        }
        if (keep) face_starts.push_back((int64_t)face_vals.size());
      }
    }
  }
  auto has = [&](const char* n) { return vert.count(n) != 0; };
  const size_t nv = vert.empty() ? 0 : vert.begin()->second.size();
  auto stack = [&](std::vector<float>* out, std::initializer_list<const char*> names) {
    for (const char* n : names)
      if (!has(n)) return false;
    const size_t w = names.size();
    out->resize(nv * w);
    size_t k = 0;
    for (const char* n : names) {
      const std::vector<float>& c = vert[n];
      for (size_t i = 0; i < nv; i++) (*out)[i * w + k] = c[i];
      k++;
    }
    return true;
  };
  stack(&sh->positions, {"x", "y", "z"});
  stack(&sh->normals, {"nx", "ny", "nz"});
  if (!vnames.empty()) {  // get_tex_coords, src/shape.jl:265-278: only the FIRST vertex property decides s,t vs u,v
    bool ok = vnames[0] == "s" ? stack(&sh->texcoords, {"s", "t"}) : stack(&sh->texcoords, {"u", "v"});
    if (ok)
      for (size_t i = 0; i < nv; i++) sh->texcoords[2 * i + 1] = 1.0f - sh->texcoords[2 * i + 1];  // flip, :233-235
  }
  if (has("alpha")) stack(&sh->colors, {"red", "green", "blue", "alpha"});  // the rgb-only path is broken upstream (:280-299)
  if (have_faces) faces_to_elements(face_vals, face_starts, sh);
}

// ---------------------------------------------------------------------------------------------------------------------
// textures, src/scene.jl:164-189
// ---------------------------------------------------------------------------------------------------------------------
struct HostTexture {
  int64_t width = 0, height = 0;
  int linear = 0;
  std::vector<float> pixelsf;    // 4 per texel
  std::vector<uint8_t> pixelsb;  // 4 per texel
};

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// 8-bit, non-interlaced PNG of any colour type -> what PIL's convert("RGBA") / convert("RGB") gives the Python mirror:
// RGBA when the file carries alpha (colour types 4 and 6, or a tRNS chunk), else RGB with the alpha byte set to 1
// (Vec4b(::RGB) stores alpha = 1, not 255: src/math.jl:39-44).
void load_png(const std::string& path, HostTexture* t) {
  std::vector<uint8_t> f = read_file(path);
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (f.size() < 33 || memcmp(f.data(), sig, 8) != 0) fail(JT_ERR_INVALID, path + ": not a PNG file");
  uint32_t w = 0, h = 0;
  int depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte, trns;
  bool have_trns = false;
  size_t pos = 8;
  while (pos + 12 <= f.size()) {
    uint32_t len = be32(f.data() + pos);
    const char* type = (const char*)f.data() + pos + 4;
    const uint8_t* body = f.data() + pos + 8;
    if (pos + 12 + (size_t)len > f.size()) fail(JT_ERR_INVALID, path + ": truncated PNG chunk");
    if (!memcmp(type, "IHDR", 4)) {
      w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
    } else if (!memcmp(type, "PLTE", 4)) {
      plte.assign(body, body + len);
    } else if (!memcmp(type, "tRNS", 4)) {
      trns.assign(body, body + len);
      have_trns = true;
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + (size_t)len;
  }
  if (w == 0 || h == 0) fail(JT_ERR_INVALID, path + ": PNG without IHDR");
  if (depth != 8 || interlace != 0)
    fail(JT_ERR_UNSUPPORTED, path + ": only 8-bit non-interlaced PNGs are supported (every texture the reference ships is one)");
  int ch;
  switch (ctype) {
    case 0: ch = 1; break;
    case 2: ch = 3; break;
    case 3: ch = 1; break;
    case 4: ch = 2; break;
    case 6: ch = 4; break;
    default: fail(JT_ERR_INVALID, path + ": bad PNG colour type");
  }
  const size_t stride = (size_t)w * (size_t)ch;
  std::vector<uint8_t> raw((stride + 1) * (size_t)h);
  uLongf out_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size())
    fail(JT_ERR_INVALID, path + ": PNG inflate failed");
  std::vector<uint8_t> img(stride * (size_t)h);
  std::vector<uint8_t> zero(stride, 0);
  for (uint32_t y = 0; y < h; y++) {  // undo the scanline filters
    const uint8_t* src = raw.data() + (stride + 1) * y;
    const int filter = src[0];
    src++;
    uint8_t* dst = img.data() + stride * y;
    const uint8_t* up = y ? dst - stride : zero.data();
    for (size_t x = 0; x < stride; x++) {
      int a = x >= (size_t)ch ? dst[x - (size_t)ch] : 0, b = up[x], c = x >= (size_t)ch ? up[x - (size_t)ch] : 0;
      int v = src[x];
      switch (filter) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: {
          int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
          v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
          break;
        }
        default: fail(JT_ERR_INVALID, path + ": bad PNG filter");
      }
      dst[x] = (uint8_t)v;
    }
  }
  const bool rgba = ctype == 4 || ctype == 6 || have_trns;
  t->width = w;
  t->height = h;
  t->linear = 0;
  t->pixelsb.resize((size_t)w * h * 4);
  for (size_t i = 0; i < (size_t)w * h; i++) {
    uint8_t r, g, b, a = 255;
    const uint8_t* p = img.data() + i * (size_t)ch;
    switch (ctype) {
      case 0:
        r = g = b = p[0];
        if (have_trns && trns.size() >= 2 && trns[1] == p[0]) a = 0;
        break;
      case 2:
        r = p[0]; g = p[1]; b = p[2];
        if (have_trns && trns.size() >= 6 && trns[1] == r && trns[3] == g && trns[5] == b) a = 0;
        break;
      case 3:
        if ((size_t)p[0] * 3 + 2 >= plte.size()) fail(JT_ERR_INVALID, path + ": PNG palette index out of range");
        r = plte[(size_t)p[0] * 3]; g = plte[(size_t)p[0] * 3 + 1]; b = plte[(size_t)p[0] * 3 + 2];
        if (have_trns && (size_t)p[0] < trns.size()) a = trns[p[0]];
        break;
      case 4: r = g = b = p[0]; a = p[1]; break;
      default: r = p[0]; g = p[1]; b = p[2]; a = p[3]; break;
    }
    uint8_t* o = t->pixelsb.data() + 4 * i;
    o[0] = r; o[1] = g; o[2] = b;
    o[3] = rgba ? a : 1;
  }
}

// Radiance RGBE -> float RGB = mantissa * 2^(e - 136) (exact), then the reference loader's rule (Q9, pinned in
// tests/test_oracle.py): display-encode, clamp to [0, 1], 16-bit quantum; alpha = 1.
void load_hdr(const std::string& path, HostTexture* t) {
  std::vector<uint8_t> f = read_file(path);
  size_t pos = 0;
  auto line = [&]() {
    std::string s;
    while (pos < f.size() && f[pos] != '\n') s += (char)f[pos++];
    if (pos < f.size()) pos++;
    return s;
  };
  std::string first = line();
  if (first.compare(0, 2, "#?") != 0) fail(JT_ERR_INVALID, path + ": not a Radiance HDR file");
  for (;;) {
    if (pos >= f.size()) fail(JT_ERR_INVALID, path + ": HDR header ends early");
    std::string s = line();
    if (s.empty() || s == "\r") break;
  }
  std::string res = line();
  int hh = 0, ww = 0;
  if (sscanf(res.c_str(), "-Y %d +X %d", &hh, &ww) != 2 || hh <= 0 || ww <= 0)
    fail(JT_ERR_UNSUPPORTED, path + ": only '-Y h +X w' HDR orientation is supported");
  std::vector<uint8_t> rgbe((size_t)ww * hh * 4);
  const uint8_t* p = f.data() + pos;
  const uint8_t* end = f.data() + f.size();
  for (int y = 0; y < hh; y++) {
    uint8_t* row = rgbe.data() + (size_t)y * ww * 4;
    if (ww >= 8 && ww <= 0x7fff && end - p >= 4 && p[0] == 2 && p[1] == 2 && !(p[2] & 0x80) && ((p[2] << 8) | p[3]) == ww) {
      p += 4;  // new-style RLE: the four channels one after the other
      for (int c = 0; c < 4; c++) {
        int x = 0;
        while (x < ww) {
          if (end - p < 2) fail(JT_ERR_INVALID, path + ": HDR data ends early");
          int n = *p++;
          if (n > 128) {
            n -= 128;
            uint8_t v = *p++;
            if (x + n > ww) fail(JT_ERR_INVALID, path + ": bad HDR run");
            for (int k = 0; k < n; k++) row[4 * (x++) + c] = v;
          } else {
            if (n == 0 || x + n > ww || end - p < n) fail(JT_ERR_INVALID, path + ": bad HDR run");
            for (int k = 0; k < n; k++) row[4 * (x++) + c] = *p++;
          }
        }
      }
    } else {  // flat pixels
      if (end - p < (ptrdiff_t)ww * 4) fail(JT_ERR_INVALID, path + ": HDR data ends early");
      memcpy(row, p, (size_t)ww * 4);
      p += (size_t)ww * 4;
    }
  }
  t->width = ww;
  t->height = hh;
  t->linear = 1;
  t->pixelsf.resize((size_t)ww * hh * 4);
  const double expo = 1.0 / 2.4;
  for (size_t i = 0; i < (size_t)ww * hh; i++) {
    const uint8_t* q = rgbe.data() + 4 * i;
    float scale = q[3] ? (float)ldexp(1.0, (int)q[3] - 136) : 0.0f;
    for (int c = 0; c < 3; c++) {
      float lin = q[3] ? (float)q[c] * scale : 0.0f;
      double v = lin > 0.0f ? (double)lin : 0.0;
      double enc = v <= 0.0031308 ? 12.92 * v : 1.055 * pow(v, expo) - 0.055;
      enc = enc < 0.0 ? 0.0 : (enc > 1.0 ? 1.0 : enc);
      t->pixelsf[4 * i + (size_t)c] = (float)(nearbyint(enc * 65535.0) / 65535.0);  // ImageMagick Q16 quantum
    }
    t->pixelsf[4 * i + 3] = 1.0f;
  }
}

void load_texture(const std::string& path, HostTexture* t) {
  size_t dot = path.rfind('.');
  std::string ext = dot == std::string::npos ? "" : path.substr(dot);
  for (char& c : ext) c = (char)tolower((unsigned char)c);
  if (ext == ".hdr") load_hdr(path, t);
  else if (ext == ".png") load_png(path, t);
  else fail(JT_ERR_UNSUPPORTED, "unknown texture format: " + ext);
}

// ---------------------------------------------------------------------------------------------------------------------
// geometry helpers in the reference's Float32 operation order
// ---------------------------------------------------------------------------------------------------------------------
float tri_area(const float* p0, const float* p1, const float* p2) {  // triangle_area, src/geometry.jl:260-262
  float a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
  float c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
  float d = (c[0] * c[0] + c[1] * c[1]) + c[2] * c[2];
  return sqrtf(d) / 2.0f;
}

}  // namespace

// =====================================================================================================================
struct jt_host_scene {
  std::vector<jt_camera> cameras;
  std::vector<std::string> camera_names;
  std::vector<jt_instance> instances;
  std::vector<jt_environment> environments;
  std::vector<jt_material> materials;
  std::vector<HostShape> shapes;
  std::vector<HostTexture> textures;
  std::vector<std::string> notes;
  // built by jt_host_scene_build
  bool built = false;
  std::vector<jt_bvh_node> tlas_nodes;
  std::vector<int64_t> tlas_prims;
  struct Light { int64_t instance, environment; std::vector<float> cdf; };
  std::vector<Light> lights;
  // the flat description
  std::vector<jt_shape_desc> shape_descs;
  std::vector<jt_texture_desc> texture_descs;
  std::vector<jt_light_desc> light_descs;
  std::vector<float> lut;
  jt_scene_desc desc;
};

namespace {

int material_type(const std::string& s) {  // src/scene.jl:201-211; unknown names fall back to matte like the Python mirror
  static const std::pair<const char*, int> names[] = {{"matte", 0}, {"glossy", 1}, {"reflective", 2}, {"transparent", 3},
                                                      {"refractive", 4}, {"subsurface", 5}, {"volume", 6},
                                                      {"volumetric", 6}, {"gltfpbr", 7}};
  for (const auto& kv : names)
    if (s == kv.first) return kv.second;
  return 0;
}

void load_scene_impl(const std::string& filename, jt_host_scene* S) {
  std::vector<uint8_t> text = read_file(filename);
  JParser jp{(const char*)text.data(), (const char*)text.data() + text.size(), filename};
  JVal js = jp.value();
  if (js.kind != JVal::Obj) fail(JT_ERR_INVALID, filename + ": top-level JSON object expected");
  std::string dir;
  size_t slash = filename.rfind('/');
  if (slash != std::string::npos) dir = filename.substr(0, slash + 1);
  static const JVal empty_arr = [] { JVal v; v.kind = JVal::Arr; return v; }();
  auto list = [&](const char* key) -> const std::vector<JVal>& {
    const JVal* v = js.get(key);
    return (v && v->kind == JVal::Arr) ? v->arr : empty_arr.arr;
  };
  for (const JVal& c : list("cameras")) {  // CameraData defaults, src/scene.jl:58-66
    no_lookat(&c, "camera");
    jt_camera cam;
    memset(&cam, 0, sizeof(cam));
    jframe(&c, &cam.frame);
    const JVal* o = c.get("orthographic");
    cam.orthographic = (o && o->kind == JVal::Bool && o->b) ? 1 : 0;
    cam.lens = jnum(&c, "lens", 0.050);
    cam.film = jnum(&c, "film", 0.036);
    cam.aspect = jnum(&c, "aspect", 1.5);
    cam.focus = jnum(&c, "focus", 10000);
    cam.aperture = jnum(&c, "aperture", 0);
    S->cameras.push_back(cam);
    const JVal* n = c.get("name");
    S->camera_names.push_back((n && n->kind == JVal::Str) ? n->str : "");
  }
  for (const JVal& t : list("textures")) {
    const JVal* uri = t.get("uri");
    if (!uri || uri->kind != JVal::Str) fail(JT_ERR_INVALID, filename + ": texture without uri");
    S->textures.emplace_back();
    if (file_exists(dir + uri->str)) {
      load_texture(dir + uri->str, &S->textures.back());
    } else {
      S->notes.push_back("missing texture " + uri->str + " -> 1x1 opaque white");
      HostTexture& w = S->textures.back();
      w.width = w.height = 1;
      w.linear = 0;
      w.pixelsb.assign(4, 255);
    }
  }
  for (const JVal& m : list("materials")) {  // MaterialData defaults, src/scene.jl:231-245
    jt_material M;
    memset(&M, 0, sizeof(M));
    const JVal* ty = m.get("type");
    M.type = material_type((ty && ty->kind == JVal::Str) ? ty->str : "matte");
    jvec3(&m, "emission", M.emission);
    jvec3(&m, "color", M.color);
    M.roughness = jnum(&m, "roughness", 0);
    M.metallic = jnum(&m, "metallic", 0);
    M.ior = jnum(&m, "ior", 1.5);
    jvec3(&m, "scattering", M.scattering);
    M.scanisotropy = jnum(&m, "scanisotropy", 0);
    M.trdepth = jnum(&m, "trdepth", 0.01);
    M.opacity = jnum(&m, "opacity", 1);
    M.emission_tex = jid(&m, "emission_tex");
    M.color_tex = jid(&m, "color_tex");
    M.roughness_tex = jid(&m, "roughness_tex");
    M.scattering_tex = jid(&m, "scattering_tex");
    M.normal_tex = jid(&m, "normal_tex");
    S->materials.push_back(M);
  }
  std::vector<char> shape_missing;
  for (const JVal& s : list("shapes")) {
    const JVal* uri = s.get("uri");
    if (!uri || uri->kind != JVal::Str) fail(JT_ERR_INVALID, filename + ": shape without uri");
    S->shapes.emplace_back();
    if (file_exists(dir + uri->str)) {
      load_shape(dir + uri->str, &S->shapes.back());
      shape_missing.push_back(0);
    } else {
      S->notes.push_back("missing shape " + uri->str + " -> its instances are dropped");
      shape_missing.push_back(1);
    }
  }
  size_t total = 0, dropped = 0;
  for (const JVal& x : list("instances")) {
    no_lookat(&x, "instance");
    jt_instance I;
    memset(&I, 0, sizeof(I));
    jframe(&x, &I.frame);
    I.shape = jid(&x, "shape");
    I.material = jid(&x, "material");
    total++;
    if (I.shape >= 1 && (size_t)I.shape <= shape_missing.size() && shape_missing[(size_t)I.shape - 1]) {
      dropped++;
      continue;
    }
    S->instances.push_back(I);
  }
  if (dropped)
    S->notes.push_back("dropped " + std::to_string(dropped) + " of " + std::to_string(total) + " instances (absent shape files)");
  for (const JVal& e : list("environments")) {
    no_lookat(&e, "environment");
    jt_environment E;
    memset(&E, 0, sizeof(E));
    jframe(&e, &E.frame);
    jvec3(&e, "emission", E.emission);
    E.emission_tex = jid(&e, "emission_tex");
    S->environments.push_back(E);
  }
}

// triangle_bounds / quad_bounds (src/geometry.jl:64-68): triangles take precedence
void shape_bboxes(const HostShape& sh, std::vector<float>* boxes) {
  const bool tris = !sh.triangles.empty();
  const std::vector<int64_t>& idx = tris ? sh.triangles : sh.quads;
  const int w = tris ? 3 : 4;
  const size_t n = idx.size() / (size_t)w, nv = sh.positions.size() / 3;
  boxes->resize(6 * n);
  for (size_t e = 0; e < n; e++) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = 0; k < w; k++) {
      int64_t v = idx[e * (size_t)w + (size_t)k];
      if (v < 1 || (size_t)v > nv) fail(JT_ERR_INVALID, "shape element " + std::to_string(e) + ": vertex id out of range");
      const float* p = sh.positions.data() + 3 * (size_t)(v - 1);
      for (int a = 0; a < 3; a++) {
        lo[a] = p[a] < lo[a] ? p[a] : lo[a];
        hi[a] = p[a] > hi[a] ? p[a] : hi[a];
      }
    }
    memcpy(boxes->data() + 6 * e, lo, 12);
    memcpy(boxes->data() + 6 * e + 3, hi, 12);
  }
}

void make_tree(const std::vector<float>& boxes, int high_quality, std::vector<jt_bvh_node>* nodes, std::vector<int64_t>* prims) {
  const int64_t n = (int64_t)boxes.size() / 6;
  nodes->assign((size_t)(2 * n + 1), jt_bvh_node());
  prims->assign((size_t)(n > 0 ? n : 1), 0);
  int64_t count = 0;
  int rc = jt_make_bvh(n ? boxes.data() : nullptr, n, high_quality, nodes->data(), &count, prims->data());
  if (rc) fail(rc, jt_last_error());
  nodes->resize((size_t)count);
  prims->resize((size_t)n);
}

void build_impl(jt_host_scene* S, int high_quality) {
  // make_scene_bvh, src/bvh.jl:66-136
  for (HostShape& sh : S->shapes) {
    std::vector<float> boxes;
    shape_bboxes(sh, &boxes);
    make_tree(boxes, high_quality, &sh.nodes, &sh.prims);
  }
  std::vector<float> iboxes(6 * S->instances.size());
  for (size_t i = 0; i < S->instances.size(); i++) {
    const jt_instance& I = S->instances[i];
    if (I.shape < 1 || (size_t)I.shape > S->shapes.size()) fail(JT_ERR_INVALID, "instance " + std::to_string(i + 1) + ": shape id out of range");
    const HostShape& sh = S->shapes[(size_t)I.shape - 1];
    if (sh.empty())
      fail(JT_ERR_INVALID, "instance " + std::to_string(i + 1) + " references an element-less shape: its Inf box hangs the reference's partition (SURVEY.md App. D); drop it first");
    const jt_bvh_node& root = sh.nodes[0];
    // transform_bbox, src/geometry.jl:70-86: min / max over the 8 transformed corners, ((x*p1 + y*p2) + z*p3) + o
    const float* x = I.frame.x; const float* y = I.frame.y; const float* z = I.frame.z; const float* o = I.frame.o;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int cx = 0; cx < 2; cx++)
      for (int cy = 0; cy < 2; cy++)
        for (int cz = 0; cz < 2; cz++) {
          float px = cx ? root.bbox_max[0] : root.bbox_min[0], py = cy ? root.bbox_max[1] : root.bbox_min[1],
                pz = cz ? root.bbox_max[2] : root.bbox_min[2];
          for (int a = 0; a < 3; a++) {
            float p = ((x[a] * px + y[a] * py) + z[a] * pz) + o[a];
            lo[a] = fminf(lo[a], p);
            hi[a] = fmaxf(hi[a], p);
          }
        }
    memcpy(iboxes.data() + 6 * i, lo, 12);
    memcpy(iboxes.data() + 6 * i + 3, hi, 12);
  }
  make_tree(iboxes, high_quality, &S->tlas_nodes, &S->tlas_prims);

  // make_trace_lights, src/trace.jl:117-187
  S->lights.clear();
  for (size_t h = 0; h < S->instances.size(); h++) {
    const jt_instance& I = S->instances[h];
    if (I.material < 1 || (size_t)I.material > S->materials.size()) fail(JT_ERR_INVALID, "instance " + std::to_string(h + 1) + ": material id out of range");
    const jt_material& M = S->materials[(size_t)I.material - 1];
    if (M.emission[0] == 0.0f && M.emission[1] == 0.0f && M.emission[2] == 0.0f) continue;
    const HostShape& sh = S->shapes[(size_t)I.shape - 1];
    if (sh.empty()) continue;
    jt_host_scene::Light L{(int64_t)h + 1, -1, {}};
    const float* P = sh.positions.data();
    if (!sh.triangles.empty()) {
      const size_t n = sh.triangles.size() / 3;
      L.cdf.resize(n);
      float acc = 0.0f;
      for (size_t e = 0; e < n; e++) {
        const int64_t* t = sh.triangles.data() + 3 * e;
        float a = tri_area(P + 3 * (t[0] - 1), P + 3 * (t[1] - 1), P + 3 * (t[2] - 1));
        acc = e ? acc + a : a;  // sequential Float32 prefix sum (:172-181)
        L.cdf[e] = acc;
      }
    }
    if (!sh.quads.empty()) {  // a second `if`, like the reference: quads overwrite
      const size_t n = sh.quads.size() / 4;
      L.cdf.resize(n);
      float acc = 0.0f;
      for (size_t e = 0; e < n; e++) {
        const int64_t* q = sh.quads.data() + 4 * e;
        float a = tri_area(P + 3 * (q[0] - 1), P + 3 * (q[1] - 1), P + 3 * (q[3] - 1)) +
                  tri_area(P + 3 * (q[2] - 1), P + 3 * (q[3] - 1), P + 3 * (q[1] - 1));
        acc = e ? acc + a : a;
        L.cdf[e] = acc;
      }
    }
    S->lights.push_back(std::move(L));
  }
  for (size_t h = 0; h < S->environments.size(); h++) {
    const jt_environment& E = S->environments[h];
    if (E.emission[0] == 0.0f && E.emission[1] == 0.0f && E.emission[2] == 0.0f) continue;
    jt_host_scene::Light L{-1, (int64_t)h + 1, {}};
    if (E.emission_tex != -1) {
      if (E.emission_tex < 1 || (size_t)E.emission_tex > S->textures.size()) fail(JT_ERR_INVALID, "environment texture id out of range");
      const HostTexture& T = S->textures[(size_t)E.emission_tex - 1];
      const size_t n = (size_t)(T.width * T.height);
      L.cdf.resize(n);
      const float pi = (float)M_PI;
      float acc = 0.0f;
      for (size_t i = 0; i < n; i++) {
        float v[4];
        if (!T.pixelsf.empty()) memcpy(v, T.pixelsf.data() + 4 * i, 16);
        else for (int c = 0; c < 4; c++) v[c] = (float)T.pixelsb[4 * i + (size_t)c] / 255.0f;
        float value = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));  // maximum over RGBA incl. alpha (Q8)
        float j = (float)(i / (size_t)T.width);
        float th = ((j + 0.5f) * pi) / (float)T.height;
        float w = value * (float)sin((double)th);
        acc = i ? acc + w : w;
        L.cdf[i] = acc;
      }
    }
    S->lights.push_back(std::move(L));
  }
  S->built = true;
}

void fill_desc(jt_host_scene* S) {
  jt_scene_desc& d = S->desc;
  memset(&d, 0, sizeof(d));
  d.num_cameras = (int64_t)S->cameras.size();       d.cameras = S->cameras.data();
  d.num_instances = (int64_t)S->instances.size();   d.instances = S->instances.data();
  d.num_environments = (int64_t)S->environments.size(); d.environments = S->environments.data();
  d.num_materials = (int64_t)S->materials.size();   d.materials = S->materials.data();
  S->shape_descs.assign(S->shapes.size(), jt_shape_desc());
  for (size_t i = 0; i < S->shapes.size(); i++) {
    const HostShape& h = S->shapes[i];
    jt_shape_desc& o = S->shape_descs[i];
    memset(&o, 0, sizeof(o));
    o.positions = h.positions.empty() ? nullptr : h.positions.data(); o.num_positions = (int64_t)h.positions.size() / 3;
    o.normals = h.normals.empty() ? nullptr : h.normals.data();       o.num_normals = (int64_t)h.normals.size() / 3;
    o.texcoords = h.texcoords.empty() ? nullptr : h.texcoords.data(); o.num_texcoords = (int64_t)h.texcoords.size() / 2;
    o.colors = h.colors.empty() ? nullptr : h.colors.data();          o.num_colors = (int64_t)h.colors.size() / 4;
    o.triangles = h.triangles.empty() ? nullptr : h.triangles.data(); o.num_triangles = (int64_t)h.triangles.size() / 3;
    o.quads = h.quads.empty() ? nullptr : h.quads.data();             o.num_quads = (int64_t)h.quads.size() / 4;
    o.bvh.nodes = h.nodes.empty() ? nullptr : h.nodes.data();         o.bvh.num_nodes = (int64_t)h.nodes.size();
    o.bvh.primitives = h.prims.empty() ? nullptr : h.prims.data();    o.bvh.num_primitives = (int64_t)h.prims.size();
  }
  d.num_shapes = (int64_t)S->shapes.size();
  d.shapes = S->shape_descs.data();
  S->texture_descs.assign(S->textures.size(), jt_texture_desc());
  for (size_t i = 0; i < S->textures.size(); i++) {
    const HostTexture& t = S->textures[i];
    jt_texture_desc& o = S->texture_descs[i];
    memset(&o, 0, sizeof(o));
    o.width = t.width; o.height = t.height; o.linear = t.linear;
    o.pixelsf = t.pixelsf.empty() ? nullptr : t.pixelsf.data();
    o.pixelsb = t.pixelsb.empty() ? nullptr : t.pixelsb.data();
  }
  d.num_textures = (int64_t)S->textures.size();
  d.textures = S->texture_descs.data();
  S->light_descs.assign(S->lights.size(), jt_light_desc());
  for (size_t i = 0; i < S->lights.size(); i++) {
    jt_light_desc& o = S->light_descs[i];
    o.instance = S->lights[i].instance; o.environment = S->lights[i].environment;
    o.elements_cdf = S->lights[i].cdf.empty() ? nullptr : S->lights[i].cdf.data();
    o.num_elements = (int64_t)S->lights[i].cdf.size();
  }
  d.num_lights = (int64_t)S->lights.size();
  d.lights = S->light_descs.data();
  d.bvh.nodes = S->tlas_nodes.empty() ? nullptr : S->tlas_nodes.data();
  d.bvh.num_nodes = (int64_t)S->tlas_nodes.size();
  d.bvh.primitives = S->tlas_prims.empty() ? nullptr : S->tlas_prims.data();
  d.bvh.num_primitives = (int64_t)S->tlas_prims.size();
  // srgb_to_rgb(b / 255f0), src/color.jl:12-23; Julia's x^2.4f0 = Float32(exp2(log2(Float64(x)) * Float64(2.4f0)))
  S->lut.resize(256);
  for (int b = 0; b < 256; b++) {
    float c = (float)b / 255.0f;
    float lo = c / 12.92f;
    float base = (c + 0.055f) / 1.055f;
    float hi = (float)exp2(log2((double)base) * (double)2.4f);
    S->lut[(size_t)b] = c <= 0.04045f ? lo : hi;
  }
  d.srgb_to_rgb_lut = S->lut.data();
}

template <class F>
int guarded(const char* who, F&& body) {
  try {
    body();
    return JT_OK;
  } catch (const Fail& e) {
    return jt_set_error(e.code, "%s: %s", who, e.what());
  } catch (const std::exception& e) {
    return jt_set_error(JT_ERR_INTERNAL, "%s: %s", who, e.what());
  } catch (...) {
    return jt_set_error(JT_ERR_INTERNAL, "%s: unknown exception", who);
  }
}

}  // namespace

extern "C" int jt_host_scene_load(const char* json_path, jt_host_scene** out) {
  if (!json_path || !out) return jt_set_error(JT_ERR_INVALID, "jt_host_scene_load: null argument");
  *out = nullptr;
  std::unique_ptr<jt_host_scene> S(new (std::nothrow) jt_host_scene());
  if (!S) return jt_set_error(JT_ERR_INTERNAL, "out of host memory");
  int rc = guarded("jt_host_scene_load", [&] {
    load_scene_impl(json_path, S.get());
    fill_desc(S.get());
  });
  if (rc) return rc;
  *out = S.release();
  return JT_OK;
}

extern "C" void jt_host_scene_destroy(jt_host_scene* S) { delete S; }

extern "C" int jt_host_scene_build(jt_host_scene* S, int high_quality_bvh) {
  if (!S) return jt_set_error(JT_ERR_INVALID, "jt_host_scene_build: null argument");
  return guarded("jt_host_scene_build", [&] {
    build_impl(S, high_quality_bvh ? 1 : 0);
    fill_desc(S);
  });
}

extern "C" int jt_host_scene_desc(jt_host_scene* S, const jt_scene_desc** out) {
  if (!S || !out) return jt_set_error(JT_ERR_INVALID, "jt_host_scene_desc: null argument");
  *out = &S->desc;
  return JT_OK;
}

// find_camera, src/scene.jl:358-370: the named camera, else "default", "camera", "camera0", "camera1", else the first
extern "C" int jt_host_scene_find_camera(jt_host_scene* S, const char* name, int32_t* camera) {
  if (!S || !camera) return jt_set_error(JT_ERR_INVALID, "jt_host_scene_find_camera: null argument");
  *camera = -1;
  if (S->cameras.empty()) return JT_OK;
  const std::string candidates[5] = {name ? name : "", "default", "camera", "camera0", "camera1"};
  for (const std::string& n : candidates)
    for (size_t i = 0; i < S->camera_names.size(); i++)
      if (S->camera_names[i] == n) {
        *camera = (int32_t)i + 1;
        return JT_OK;
      }
  *camera = 1;
  return JT_OK;
}

extern "C" int jt_host_scene_num_notes(jt_host_scene* S) { return S ? (int)S->notes.size() : 0; }
extern "C" const char* jt_host_scene_note(jt_host_scene* S, int i) {
  if (!S || i < 0 || i >= (int)S->notes.size()) return "";
  return S->notes[(size_t)i].c_str();
}
