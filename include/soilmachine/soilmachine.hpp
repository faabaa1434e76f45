// include/soilmachine/soilmachine.hpp -- C++ facade: the reference's own class names and call shapes
// for the particle/terrain hot path, forwarding to the C ABI (include/soilmachine_b200.h).
//
// The reference has no plugin boundary; its frame loop (SoilMachine.cpp:283-329) is written against
// these types, which upstream defines in source/layermap.h, source/surface.h and
// source/particle/{particle,water,wind}.h.  A maintainer swaps those four #includes for this header
// (INTEGRATION.md shows the diff) and the loop runs on the GPU:
//
//   Layermap map(SEED, ivec2(SIZEX, SIZEY), vertexpool);      // SoilMachine.cpp:83   -> sm_initialize
//   WaterParticle::run(map, vertexpool, NWATER);              // replaces :288-298    -> sm_water_run
//   WindParticle::run(map, vertexpool, NWIND);                // replaces :304-307    -> sm_wind_run
//   WaterParticle::mapfrequency(map); ...resetfrequency(map); // :313-320             -> sm_frequency_update
//
// Legacy per-cell calls (map.height(p), map.add(p, map.pool.get(h, type)), map.remove, map.surface,
// map.normal, Particle::cascade) keep working: each forwards to one sm_cell_* call (one tiny kernel,
// fine for the GUI / initialisation code that uses them, not meant for inner loops).
//
// Everything is header-only and written from scratch; vector types are minimal stand-ins that
// convert implicitly from/to any type with .x/.y(/.z) members (so glm::ivec2 etc. can be passed).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include "../soilmachine_b200.h"
#include "soilfile.hpp"

namespace soilmachine {

struct ivec2 {
  int x = 0, y = 0;
  ivec2() {}
  ivec2(int a, int b) : x(a), y(b) {}
  template <class V, class = decltype(std::declval<V>().x + std::declval<V>().y)>
  ivec2(const V& v) : x((int)v.x), y((int)v.y) {}
};
struct vec2 {
  float x = 0, y = 0;
  vec2() {}
  vec2(float a, float b) : x(a), y(b) {}
  template <class V, class = decltype(std::declval<V>().x + std::declval<V>().y)>
  vec2(const V& v) : x((float)v.x), y((float)v.y) {}
};
struct vec3 { float x = 0, y = 0, z = 0; };
struct vec4 { float x = 0, y = 0, z = 0, w = 0; };

using SurfType = size_t;

// surface.h:11-39
struct SurfParam {
  std::string name;
  float density = 0.0f, porosity = 0.0f;
  vec4 color{0.5f, 0.5f, 0.5f, 1.0f}, phong{0.5f, 0.8f, 0.2f, 32.0f};
  SurfType transports = 0; float solubility = 1.0f, equrate = 1.0f, friction = 1.0f;
  SurfType erodes = 0; float erosionrate = 0.0f;
  SurfType cascades = 0; float maxdiff = 1.0f, settling = 0.0f;
  SurfType abrades = 0; float suspension = 0.0f, abrasion = 0.0f;
};
// surface.h:65-101 (parameters only; the noise itself runs on the device, sm_initialize)
struct SurfLayer {
  SurfType type = 0;
  float min = 0.0f, bias = 0.0f, scale = 1.0f, octaves = 1.0f, lacunarity = 1.0f, gain = 0.0f, frequency = 1.0f;
  explicit SurfLayer(SurfType t) : type(t) {}
};

// the reference's global tables (surface.h:41-57,104)
inline std::vector<SurfParam>& soils_table() {
  static std::vector<SurfParam> t = [] { SurfParam air; air.name = "Air"; air.porosity = 1.0f;
    air.color = vec4{0.0f, 0.2f, 0.4f, 1.0f}; air.solubility = 0.0f; air.equrate = 0.0f; air.friction = 0.0f;
    air.maxdiff = 0.0f; return std::vector<SurfParam>{air}; }();
  return t;
}
inline std::map<std::string, int>& soilmap_table() { static std::map<std::string, int> m{{"Air", 0}}; return m; }
inline std::vector<SurfLayer>& layers_table() { static std::vector<SurfLayer> l; return l; }
// loadsoil(), io.h:7-230: fills the soils / soilmap / layers tables from a `.soil` file (same quirks as
// upstream, see soilfile.hpp) and returns the WORLD block; an application that keeps the reference's global
// ints assigns them from the result (SIZEX = w.sizex; ...).
inline WorldEntry loadsoil(const std::string& file = "soil/default.soil") {
  SoilFile f;
  parse_soil_file(file, f);
  soils_table().clear(); soilmap_table().clear(); layers_table().clear();
  for (const SoilEntry& e : f.soils) {
    SurfParam p;
    p.name = e.name; p.density = e.density; p.porosity = e.porosity;
    p.color = vec4{e.color[0], e.color[1], e.color[2], e.color[3]};
    p.phong = vec4{e.phong[0], e.phong[1], e.phong[2], e.phong[3]};
    p.transports = (SurfType)e.transports; p.solubility = e.solubility; p.equrate = e.equrate; p.friction = e.friction;
    p.erodes = (SurfType)e.erodes; p.erosionrate = e.erosionrate;
    p.cascades = (SurfType)e.cascades; p.maxdiff = e.maxdiff; p.settling = e.settling;
    p.abrades = (SurfType)e.abrades; p.suspension = e.suspension; p.abrasion = e.abrasion;
    soils_table().push_back(p);
  }
  for (const auto& kv : f.soilmap) soilmap_table()[kv.first] = kv.second;
  for (const LayerEntry& l : f.layers) {
    SurfLayer L((SurfType)l.type);
    L.min = l.min; L.bias = l.bias; L.scale = l.scale; L.octaves = l.octaves; L.lacunarity = l.lacunarity;
    L.gain = l.gain; L.frequency = l.frequency;
    layers_table().push_back(L);
  }
  return f.world;
}

#define soils (::soilmachine::soils_table())
#define soilmap (::soilmachine::soilmap_table())
#define layers (::soilmachine::layers_table())

struct Error : std::runtime_error { int code; Error(int c, const std::string& m) : std::runtime_error(m), code(c) {} };

// layermap.h:37-62 -- a value carrier on the host (the device owns the real columns)
struct sec {
  sec* next = nullptr; sec* prev = nullptr;
  SurfType type = 0; double size = 0.0, floor = 0.0, saturation = 0.0;
  sec() {}
  sec(double s, SurfType t) : type(t), size(s) {}
  void reset() { next = prev = nullptr; type = 0; size = floor = saturation = 0.0; }
};
// layermap.h:64-119
class secpool {
 public:
  int size = 0; sec* start = nullptr; std::deque<sec*> free;
  secpool() {}
  ~secpool() { delete[] start; }
  void reserve(int N) { delete[] start; start = new sec[N]; free.clear(); for (int i = 0; i < N; i++) free.push_front(start + i); size = N; }
  template <class... A> sec* get(A&&... a) {
    if (free.empty()) return nullptr;                       // layermap.h:92-95
    sec* E = free.back(); *E = sec(std::forward<A>(a)...); free.pop_back(); return E;
  }
  void unget(sec* E) { if (!E) return; E->reset(); free.push_front(E); }
  void reset() { free.clear(); for (int i = 0; i < size; i++) free.push_front(start + i); }
};

// layermap.h:127-228
class Layermap {
 public:
  ivec2 dim; secpool pool; unsigned* section = nullptr;
  sm_context* ctx = nullptr;

  Layermap(int SEED, ivec2 _dim, int SCALE = 80, int device = 0) { pool.reserve(256); open(_dim, SCALE, device); initialize(SEED, _dim); }
  template <class VP> Layermap(int SEED, ivec2 _dim, VP&, int SCALE = 80, int device = 0) : Layermap(SEED, _dim, SCALE, device) {}
  ~Layermap() { if (ctx) sm_destroy(ctx); }
  Layermap(const Layermap&) = delete;

  void initialize(int SEED, ivec2 _dim) {                   // layermap.h:163-216
    if (_dim.x != dim.x || _dim.y != dim.y) throw Error(SM_ERR_INVALID, "Layermap::initialize: size is fixed at construction");
    push_tables();
    std::vector<sm_layer> L;
    for (auto& l : layers) L.push_back(sm_layer{(int32_t)l.type, l.min, l.bias, l.scale, l.octaves, l.lacunarity, l.gain, l.frequency});
    ck(sm_initialize(ctx, SEED, L.data(), (int32_t)L.size()));
  }
  double height(ivec2 p) { double h; ck(sm_cell_query(ctx, p.x, p.y, &h, nullptr, nullptr)); return h; }   // :422
  double height(vec2 p) { double h; ck(sm_height_bilinear(ctx, p.x, p.y, &h)); return h; }                  // :427
  vec3 normal(ivec2 p) { float n[3]; ck(sm_cell_query(ctx, p.x, p.y, nullptr, nullptr, n)); return vec3{n[0], n[1], n[2]}; }  // :341
  template <class VP> vec3 normal(ivec2 p, VP&) { return normal(p); }
  SurfType surface(ivec2 p) { int32_t s; ck(sm_cell_query(ctx, p.x, p.y, nullptr, &s, nullptr)); return (SurfType)s; }       // :417
  void add(ivec2 p, sec* E) {                               // :230 (E is consumed, as upstream)
    if (!E) return;
    ck(sm_cell_add(ctx, p.x, p.y, E->size, (int32_t)E->type));
    pool.unget(E);
  }
  double remove(ivec2 p, double h) { double d; ck(sm_cell_remove(ctx, p.x, p.y, h, &d)); return d; }        // :310
  // meshing belongs to the renderer (layermap.h:443-555): accepted and ignored
  template <class VP> void meshpool(VP&) {}
  template <class VP> void update(ivec2, VP&) {}
  template <class VP> void update(VP&) {}
  template <class VP> void slice(VP&, double = 0) {}

  // A full section pool is not fatal upstream: secpool::get prints and returns NULL, add() drops the section
  // (layermap.h:92-95,232-234) and the program keeps running.  Same here: the drop is reported, the call
  // counts it in stats.pool_drops, and the status is per call.
  void ck(int rc) const {
    if (rc == SM_ERR_POOL) { std::fprintf(stderr, "Memory Pool Out-Of-Elements (%s)\n", sm_last_error(ctx)); return; }
    if (rc != SM_OK) throw Error(rc, sm_last_error(ctx));
  }
  void push_tables() {
    std::vector<sm_soil> t;
    for (auto& s : soils) t.push_back(sm_soil{(int32_t)s.transports, (int32_t)s.erodes, (int32_t)s.cascades, (int32_t)s.abrades,
                                              s.density, s.porosity, s.solubility, s.equrate, s.friction, s.erosionrate,
                                              s.maxdiff, s.settling, s.suspension, s.abrasion});
    ck(sm_set_soils(ctx, t.data(), (int32_t)t.size()));
  }
 private:
  void open(ivec2 _dim, int SCALE, int device) {
    dim = _dim;
    sm_config cfg{dim.x, dim.y, SCALE, device, 0, 0, 0};
    int rc = sm_create(&cfg, &ctx);
    if (rc != SM_OK) throw Error(rc, sm_last_error(nullptr));
  }
};

// particle.h:11-103
struct Particle {
  vec2 pos; vec2 speed; bool isalive = true;
  template <class VP> static void cascade(vec2 p, Layermap& map, VP&, int transferloop = 0) {               // particle.h:24
    map.ck(sm_cell_cascade(map.ctx, p.x, p.y, transferloop));
  }
};

namespace detail {
// spawn positions exactly as the particle constructors draw them (water.h:13, wind.h:15: GCC evaluates
// the two rand() arguments right to left, so y takes the first draw)
inline std::vector<float> spawn(const Layermap& map, int n) {
  std::vector<float> xy((size_t)n * 2);
  for (int i = 0; i < n; i++) { int y = rand() % map.dim.y; int x = rand() % map.dim.x; xy[2 * i] = (float)x; xy[2 * i + 1] = (float)y; }
  return xy;
}
}  // namespace detail

// water.h:9-373 -- the batch entry point replaces the loop SoilMachine.cpp:288-298
struct WaterParticle : Particle {
  // Host mirrors of the device maps, indexed [y*dim.x + x] as upstream (water.h:345-346).  They exist only
  // after init(dimx, dimy) (upstream: init() allocates them, water.h:21-24) and are refreshed by
  // mapfrequency() / resetfrequency(), i.e. exactly where the reference frame loop reads them
  // (SoilMachine.cpp:314-319).  A headless loop that never calls init(...) pays no download.
  inline static float* frequency = nullptr;
  inline static float* track = nullptr;
  inline static double volumeFactor = 0.015;                  // water.h:368 (fixed on the device; see flood())
  static void init() {}                                       // maps live on the device; no host mirror
  static void init(int dimx, int dimy) {
    delete[] frequency; delete[] track;
    frequency = new float[(size_t)dimx * dimy]();
    track = new float[(size_t)dimx * dimy]();
  }
  template <class VP> static sm_stats run(Layermap& map, VP&, int NWATER) {
    map.push_tables();                                        // upstream reads soils[] live (GUI sliders, :165-200)
    std::vector<float> xy = detail::spawn(map, NWATER);
    sm_stats st{};
    map.ck(sm_water_run(map.ctx, NWATER, xy.data(), 0, &st));
    return st;
  }
  // The flood tail of the per-particle loop (SoilMachine.cpp:292-296, water.h:123-145) for the whole batch:
  // every finished particle of the last run() floods, in ascending particle index.
  template <class VP> static sm_hydro_stats flood(Layermap& map, VP&) {
    if (volumeFactor != 0.015) throw Error(SM_ERR_INVALID, "WaterParticle::volumeFactor is fixed at 0.015 on the device");
    sm_hydro_stats st{};
    map.ck(sm_water_flood(map.ctx, &st));
    return st;
  }
  // WaterParticle::seep(map, vertexpool), water.h:335-343 / SoilMachine.cpp:300-301
  template <class VP> static sm_hydro_stats seep(Layermap& map, VP&) {
    if (volumeFactor != 0.015) throw Error(SM_ERR_INVALID, "WaterParticle::volumeFactor is fixed at 0.015 on the device");
    sm_hydro_stats st{};
    map.ck(sm_seep(map.ctx, &st));
    return st;
  }
  // water.h:358-365 + 353-356 fused on the device: frequency <- blend(track), track <- 0
  static void mapfrequency(Layermap& map) {
    map.ck(sm_frequency_update(map.ctx));
    if (frequency) map.ck(sm_get_frequency(map.ctx, frequency, track, nullptr));
  }
  static void resetfrequency(Layermap&) {}                                            // water.h:353-356 (done above)
  static std::vector<float> download_frequency(Layermap& map) {                       // water.h:345
    std::vector<float> f((size_t)map.dim.x * map.dim.y);
    map.ck(sm_get_frequency(map.ctx, f.data(), nullptr, nullptr));
    return f;
  }
};
// wind.h:11-140 -- the batch entry point replaces the loop SoilMachine.cpp:304-307
struct WindParticle : Particle {
  inline static float* frequency = nullptr;                   // wind.h:48; refreshed at the end of run()
  static void init() {}
  static void init(int dimx, int dimy) { delete[] frequency; frequency = new float[(size_t)dimx * dimy](); }
  template <class VP> static sm_stats run(Layermap& map, VP&, int NWIND) {
    map.push_tables();
    std::vector<float> xy = detail::spawn(map, NWIND);
    sm_stats st{};
    map.ck(sm_wind_run(map.ctx, NWIND, xy.data(), 0, &st));
    if (frequency) map.ck(sm_get_frequency(map.ctx, nullptr, nullptr, frequency));
    return st;
  }
  static std::vector<float> download_frequency(Layermap& map) {                       // wind.h:48
    std::vector<float> f((size_t)map.dim.x * map.dim.y);
    map.ck(sm_get_frequency(map.ctx, nullptr, nullptr, f.data()));
    return f;
  }
};

}  // namespace soilmachine
