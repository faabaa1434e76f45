// oracle/_ref/libsmref.so -- the REFERENCE's own hot path behind a small C API.
//
// TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; the product never does.
//
// The reference headers are #included VERBATIM from where they lie (the Makefile passes
// -I$(REFERENCE_ROOT)); nothing of the reference is copied into this repository.  What is
// restated here is only the ~40 GL-free lines of SoilMachine.cpp's main():
//   globals                      SoilMachine.cpp:9-17
//   include order                SoilMachine.cpp:19-26   (vertexpool.h -> stubs.h, scene.h dropped)
//   srand/loadsoil/init          SoilMachine.cpp:36-48
//   Vertexpool + Layermap        SoilMachine.cpp:82-83
//   frame loop                   SoilMachine.cpp:287-307, 313-320
// plus the LOCKSTEP driver, which calls the reference's own move()/interact() sweep by sweep in
// particle-index order (the canonical order the CUDA path reproduces bit for bit).
#include <iostream>
#include <fstream>
#include <sstream>
#include <deque>
#include <vector>
#include <string>
#include <functional>
#include <algorithm>
#include <map>
#include <random>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <memory>

using namespace std;

// ---- globals of SoilMachine.cpp:9-17 -------------------------------------------------------
int SIZEX = 256;
int SIZEY = 256;
int SCALE = 80;
int SLICE = 2 * SCALE;
int NWIND = 250;
int NWATER = 250;
static int g_poolsize = 10000000;  // reference: #define POOLSIZE 10000000 (made run-time sized)
#define POOLSIZE g_poolsize
int SEED;

#include <glm/glm.hpp>
using glm::uint;
#include "stubs.h"

// ---- the reference, verbatim ---------------------------------------------------------------
#include "source/layermap.h"
#include "source/particle/water.h"
#include "source/particle/wind.h"
#include "source/io.h"

// ---------------------------------------------------------------------------------------------
namespace {

struct Ctx {
  unique_ptr<Vertexpool<Vertex>> vp;
  unique_ptr<Layermap> lmap;
  vector<SurfParam> soils0;  // pristine tables (surface.h:41-57) for re-initialisation
  map<string, int> soilmap0;
  bool saved = false;
  // lockstep batches
  vector<unique_ptr<WaterParticle>> water;
  vector<int> water_live;
  vector<unique_ptr<WindParticle>> wind;
  vector<int> wind_live;
} g;

struct SoilPOD {  // numeric mirror of SurfParam (surface.h:11-39); must match tests/_refapi.py
  char name[32];
  int32_t transports, erodes, cascades, abrades;
  float density, porosity, solubility, equrate, friction, erosionrate, maxdiff, settling,
      suspension, abrasion;
  float color[4];
};

struct LayerPOD {  // numeric mirror of SurfLayer (surface.h:65-101)
  int32_t type;
  float min, bias, scale, octaves, lacunarity, gain, frequency;
};

struct Stats {          // all counters are per call
  int64_t steps;        // particle-steps = move() returned true and interact() ran
  int64_t sweeps;       // lockstep only
  int64_t exit_oob;     // water: move() false with volume == 0   | wind: move() false
  int64_t exit_evap;    // water: interact() returned false
  int64_t exit_stall;   // water: move() false with volume > 0 (flood candidate)
  double seconds;       // steady_clock around the particle loop(s)
};

inline double now() {
  return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count();
}

// restates the ctor body water.h:13-17 for an explicit spawn position
inline void respawn(WaterParticle& p, Layermap& map, float x, float y) {
  p.pos = vec2(x, y);
  p.ipos = round(p.pos);
  p.surface = map.surface(p.ipos);
  p.param = soils[p.surface];
  p.contains = p.param.transports;
}
// restates the ctor body wind.h:15-20
inline void respawn(WindParticle& p, Layermap& map, float x, float y) {
  p.pos = vec2(x, y);
  p.ipos = round(p.pos);
  p.surface = map.surface(p.ipos);
  p.param = soils[p.surface];
  p.contains = p.param.transports;
}

}  // namespace

extern "C" {

// Load a .soil file and build the terrain exactly as main() does (SoilMachine.cpp:36-48,82-83).
// dimx/dimy > 0 override the WORLD block (the BASELINE configs do this).  Returns 0 on success.
int smref_init(const char* soilfile, int seed, int dimx, int dimy, int poolsize, int quiet) {
  if (!g.saved) {
    g.soils0 = soils;
    g.soilmap0 = soilmap;
    g.saved = true;
  }
  g.water.clear(); g.wind.clear(); g.lmap.reset(); g.vp.reset();
  soils = g.soils0; soilmap = g.soilmap0; layers.clear(); phong.clear();
  SIZEX = 256; SIZEY = 256; SCALE = 80; NWIND = 250; NWATER = 250;
  {
    ifstream probe(soilfile);
    if (!probe.is_open()) return 1;  // loadsoil() would exit(0)
  }
  streambuf* old = nullptr;
  ostringstream sink;
  if (quiet) old = cout.rdbuf(sink.rdbuf());
  SEED = seed;
  srand(SEED);
  loadsoil(soilfile);
  if (dimx > 0) SIZEX = dimx;
  if (dimy > 0) SIZEY = dimy;
  SLICE = 2 * SCALE;
  g_poolsize = poolsize > 0 ? poolsize : 10000000;
  delete[] WaterParticle::frequency; delete[] WaterParticle::track; delete[] WindParticle::frequency;
  WaterParticle::init();
  WindParticle::init();
  g.vp.reset(new Vertexpool<Vertex>(SIZEX * SIZEY, 1));
  g.lmap.reset(new Layermap(SEED, ivec2(SIZEX, SIZEY), *g.vp));
  if (quiet) cout.rdbuf(old);
  return 0;
}

void smref_world(int* out5) {
  out5[0] = SIZEX; out5[1] = SIZEY; out5[2] = SCALE; out5[3] = NWATER; out5[4] = NWIND;
}
int smref_nsoils() { return (int)soils.size(); }
int smref_nlayers() { return (int)layers.size(); }
void smref_get_soils(SoilPOD* out) {
  for (size_t i = 0; i < soils.size(); i++) {
    const SurfParam& s = soils[i];
    memset(&out[i], 0, sizeof(SoilPOD));
    strncpy(out[i].name, s.name.c_str(), 31);
    out[i].transports = (int)s.transports; out[i].erodes = (int)s.erodes;
    out[i].cascades = (int)s.cascades; out[i].abrades = (int)s.abrades;
    out[i].density = s.density; out[i].porosity = s.porosity; out[i].solubility = s.solubility;
    out[i].equrate = s.equrate; out[i].friction = s.friction; out[i].erosionrate = s.erosionrate;
    out[i].maxdiff = s.maxdiff; out[i].settling = s.settling; out[i].suspension = s.suspension;
    out[i].abrasion = s.abrasion;
    out[i].color[0] = s.color.x; out[i].color[1] = s.color.y; out[i].color[2] = s.color.z;
    out[i].color[3] = s.color.w;
  }
}
void smref_get_layers(LayerPOD* out) {
  for (size_t i = 0; i < layers.size(); i++) {
    out[i].type = (int)layers[i].type; out[i].min = layers[i].min; out[i].bias = layers[i].bias;
    out[i].scale = layers[i].scale; out[i].octaves = layers[i].octaves;
    out[i].lacunarity = layers[i].lacunarity; out[i].gain = layers[i].gain;
    out[i].frequency = layers[i].frequency;
  }
}

// ---- column access (bottom -> top CSR, cell order x*dim.y + y as in layermap.h:151) -----------
int64_t smref_nsections() {
  int64_t n = 0;
  for (int x = 0; x < SIZEX; x++)
    for (int y = 0; y < SIZEY; y++)
      for (sec* s = g.lmap->top(ivec2(x, y)); s != NULL; s = s->prev) n++;
  return n;
}
int64_t smref_pool_free() { return (int64_t)g.lmap->pool.free.size(); }

void smref_get_columns(int64_t* offsets, int32_t* type, double* size, double* floor_, double* saturation) {
  int64_t n = 0;
  for (int x = 0; x < SIZEX; x++)
    for (int y = 0; y < SIZEY; y++) {
      offsets[(int64_t)x * SIZEY + y] = n;
      sec* s = g.lmap->top(ivec2(x, y));
      sec* bottom = NULL;
      for (; s != NULL; s = s->prev) bottom = s;
      // walk up via the column: prev pointers only are reliable (next is not maintained on pop),
      // so collect top->bottom then reverse
      int64_t start = n;
      for (s = g.lmap->top(ivec2(x, y)); s != NULL; s = s->prev) {
        type[n] = (int32_t)s->type; size[n] = s->size; floor_[n] = s->floor; saturation[n] = s->saturation;
        n++;
      }
      reverse(type + start, type + n); reverse(size + start, size + n);
      reverse(floor_ + start, floor_ + n); reverse(saturation + start, saturation + n);
      (void)bottom;
    }
  offsets[(int64_t)SIZEX * SIZEY] = n;
}

// Replace the terrain by explicit columns (for KATs on hand-made columns).  Sections are pushed
// with the reference's own add(); floor is whatever add() computes (layermap.h:304).
void smref_set_columns(const int64_t* offsets, const int32_t* type, const double* size,
                       const double* saturation) {
  Layermap& map = *g.lmap;
  for (int x = 0; x < SIZEX; x++)
    for (int y = 0; y < SIZEY; y++) {
      ivec2 p(x, y);
      // pop everything
      while (map.top(p) != NULL) {
        map.top(p)->size = 0.0;
        map.remove(p, 0.0);
      }
      int64_t c = (int64_t)x * SIZEY + y;
      for (int64_t k = offsets[c]; k < offsets[c + 1]; k++) {
        map.add(p, map.pool.get(size[k], (SurfType)type[k]));
        if (map.top(p) != NULL) map.top(p)->saturation = saturation ? saturation[k] : 0.0;
      }
    }
}

void smref_heights(double* out) {
  for (int x = 0; x < SIZEX; x++)
    for (int y = 0; y < SIZEY; y++) out[(int64_t)x * SIZEY + y] = g.lmap->height(ivec2(x, y));
}
void smref_surfaces(int32_t* out) {
  for (int x = 0; x < SIZEX; x++)
    for (int y = 0; y < SIZEY; y++) out[(int64_t)x * SIZEY + y] = (int32_t)g.lmap->surface(ivec2(x, y));
}

void smref_get_frequency(float* wfreq, float* wtrack, float* windfreq) {
  size_t n = (size_t)SIZEX * SIZEY;
  if (wfreq) memcpy(wfreq, WaterParticle::frequency, n * 4);
  if (wtrack) memcpy(wtrack, WaterParticle::track, n * 4);
  if (windfreq) memcpy(windfreq, WindParticle::frequency, n * 4);
}
void smref_set_frequency(const float* wfreq, const float* wtrack, const float* windfreq) {
  size_t n = (size_t)SIZEX * SIZEY;
  if (wfreq) memcpy(WaterParticle::frequency, wfreq, n * 4);
  if (wtrack) memcpy(WaterParticle::track, wtrack, n * 4);
  if (windfreq) memcpy(WindParticle::frequency, windfreq, n * 4);
}
// SoilMachine.cpp:313-320 without the texture upload
void smref_frequency_update() {
  WaterParticle::mapfrequency(*g.lmap);
  WaterParticle::resetfrequency(*g.lmap);
}

// Layermap::update(Vertexpool&) (layermap.h:551-555) into the stub vertex buffer; 11 floats per cell
void smref_mesh(int slice, float* out) {
  SLICE = slice;
  g.lmap->update(*g.vp);
  for (size_t i = 0; i < g.vp->buf.size(); i++) {
    const Vertex& v = g.vp->buf[i];
    float* o = out + 11 * i;
    o[0] = v.position[0]; o[1] = v.position[1]; o[2] = v.position[2];
    o[3] = v.normal[0]; o[4] = v.normal[1]; o[5] = v.normal[2];
    o[6] = v.color[0]; o[7] = v.color[1]; o[8] = v.color[2]; o[9] = v.color[3]; o[10] = v.index;
  }
}
// the per-pixel values exportheight / exportcolor compute (io.h:236-240, 247-250)
void smref_export(float* height, float* bgra) {
  for (int x = 0; x < SIZEX; x++) for (int y = 0; y < SIZEY; y++) {
    Vertex* v = g.vp->get(g.lmap->section, x * SIZEY + y);
    size_t i = (size_t)x * SIZEY + y;
    if (height) { vec4 c = vec4(v->position[1]/SCALE/sqrt(2), v->position[1]/SCALE/sqrt(2), v->position[1]/SCALE/sqrt(2), 1); height[i] = c.x; }
    if (bgra) { vec4 color = vec4(v->color[2], v->color[1], v->color[0], 1); bgra[4*i] = color.x; bgra[4*i+1] = color.y; bgra[4*i+2] = color.z; bgra[4*i+3] = color.w; }
  }
}

// ---- single-call KAT entry points onto the reference's Layermap ---------------------------------
double smref_height_i(int x, int y) { return g.lmap->height(ivec2(x, y)); }
double smref_height_f(float x, float y) { return g.lmap->height(vec2(x, y)); }
int smref_surface(int x, int y) { return (int)g.lmap->surface(ivec2(x, y)); }
void smref_normal(int x, int y, float* out3) {
  vec3 n = g.lmap->normal(ivec2(x, y));
  out3[0] = n.x; out3[1] = n.y; out3[2] = n.z;
}
void smref_add(int x, int y, double size, int type) {
  g.lmap->add(ivec2(x, y), g.lmap->pool.get(size, (SurfType)type));
}
double smref_remove(int x, int y, double h) { return g.lmap->remove(ivec2(x, y), h); }
void smref_cascade(float x, float y, int transferloop) {
  Particle::cascade(vec2(x, y), *g.lmap, *g.vp, transferloop);
}

// ---- spawn lists: the ctor's two rand() draws (water.h:13 / wind.h:15) --------------------------
// GCC evaluates the two ctor arguments right to left, so y takes the first draw; constructing a
// real particle here keeps whatever order this compiler picks.
void smref_srand(unsigned s) { srand(s); }
void smref_spawn_list(int n, float* xy) {
  for (int i = 0; i < n; i++) {
    WaterParticle p(*g.lmap);
    xy[2 * i] = p.pos.x; xy[2 * i + 1] = p.pos.y;
  }
}

// ---- LOCKSTEP water ---------------------------------------------------------------------------------
void smref_water_begin(int n, const float* xy) {
  g.water.clear(); g.water_live.clear();
  for (int i = 0; i < n; i++) {
    g.water.emplace_back(new WaterParticle(*g.lmap));
    respawn(*g.water.back(), *g.lmap, xy[2 * i], xy[2 * i + 1]);
    g.water_live.push_back(i);
  }
}
// one sweep: every live particle, ascending index, does move() && interact()
int smref_water_sweep(Stats* st) {
  vector<int> next;
  next.reserve(g.water_live.size());
  for (int i : g.water_live) {
    WaterParticle& p = *g.water[i];
    if (!p.move(*g.lmap, *g.vp)) {
      if (p.volume == 0.0) st->exit_oob++; else st->exit_stall++;
      continue;
    }
    st->steps++;
    if (!p.interact(*g.lmap, *g.vp)) { st->exit_evap++; continue; }
    next.push_back(i);
  }
  g.water_live.swap(next);
  st->sweeps++;
  return (int)g.water_live.size();
}
void smref_water_state(float* pos, float* speed, double* volume, double* sediment, int32_t* contains,
                       int32_t* alive) {
  size_t n = g.water.size();
  for (size_t i = 0; i < n; i++) alive[i] = 0;
  for (int i : g.water_live) alive[i] = 1;
  for (size_t i = 0; i < n; i++) {
    WaterParticle& p = *g.water[i];
    pos[2 * i] = p.pos.x; pos[2 * i + 1] = p.pos.y;
    speed[2 * i] = p.Particle::speed.x; speed[2 * i + 1] = p.Particle::speed.y;
    volume[i] = p.volume; sediment[i] = p.sediment; contains[i] = (int32_t)p.contains;
  }
}
void smref_water_run(int n, const float* xy, int max_sweeps, Stats* st) {
  memset(st, 0, sizeof(Stats));
  smref_water_begin(n, xy);
  double t0 = now();
  while (!g.water_live.empty() && (max_sweeps <= 0 || st->sweeps < max_sweeps)) smref_water_sweep(st);
  st->seconds = now() - t0;
}

// ---- pooling hydrology after a lockstep batch ----------------------------
// Every particle of the finished water batch that is no longer live gets its flood() call
// (water.h:123-145), in ascending particle index.  A flood is atomic: the nested particles
// WaterParticle::cascade (water.h:151-283) spawns run to completion inside it, as upstream.  flood()
// itself decides who floods (volume >= minvol and spill left).  Returns how many did.
// WaterParticle::volumeFactor is a mutable static upstream (water.h:33,368)
void smref_set_volume_factor(double v) { WaterParticle::volumeFactor = v; }
int64_t smref_water_flood(void) {
  vector<char> live(g.water.size(), 0);
  for (int i : g.water_live) live[i] = 1;
  int64_t floods = 0;
  for (size_t i = 0; i < g.water.size(); i++) {
    if (live[i]) continue;
    WaterParticle& p = *g.water[i];
    if (!(p.volume < p.minvol) && p.spill > 0) floods++;
    p.flood(*g.lmap, *g.vp);
  }
  return floods;
}
// the per-frame full-grid pass WaterParticle::seep(map, vertexpool), water.h:335-343 / SoilMachine.cpp:300-301
void smref_seep(void) { WaterParticle::seep(*g.lmap, *g.vp); }

// ---- LOCKSTEP wind ----------------------------------------------------------------------------------
void smref_wind_begin(int n, const float* xy) {
  g.wind.clear(); g.wind_live.clear();
  for (int i = 0; i < n; i++) {
    g.wind.emplace_back(new WindParticle(*g.lmap));
    respawn(*g.wind.back(), *g.lmap, xy[2 * i], xy[2 * i + 1]);
    g.wind_live.push_back(i);
  }
}
int smref_wind_sweep(Stats* st) {
  vector<int> next;
  next.reserve(g.wind_live.size());
  for (int i : g.wind_live) {
    WindParticle& p = *g.wind[i];
    if (!p.move(*g.lmap, *g.vp)) { st->exit_oob++; continue; }
    st->steps++;
    if (!p.interact(*g.lmap, *g.vp)) { st->exit_evap++; continue; }
    next.push_back(i);
  }
  g.wind_live.swap(next);
  st->sweeps++;
  return (int)g.wind_live.size();
}
void smref_wind_state(float* pos, float* speed3, double* height, double* sediment, int32_t* contains,
                      int32_t* alive) {
  size_t n = g.wind.size();
  for (size_t i = 0; i < n; i++) alive[i] = 0;
  for (int i : g.wind_live) alive[i] = 1;
  for (size_t i = 0; i < n; i++) {
    WindParticle& p = *g.wind[i];
    pos[2 * i] = p.pos.x; pos[2 * i + 1] = p.pos.y;
    speed3[3 * i] = p.speed.x; speed3[3 * i + 1] = p.speed.y; speed3[3 * i + 2] = p.speed.z;
    height[i] = p.height; sediment[i] = p.sediment; contains[i] = (int32_t)p.contains;
  }
}
void smref_wind_run(int n, const float* xy, int max_sweeps, Stats* st) {
  memset(st, 0, sizeof(Stats));
  smref_wind_begin(n, xy);
  double t0 = now();
  while (!g.wind_live.empty() && (max_sweeps <= 0 || st->sweeps < max_sweeps)) smref_wind_sweep(st);
  st->seconds = now() - t0;
}

// ---- SEQUENTIAL reference loops (SoilMachine.cpp:287-307), explicit spawn list or rand() --------
// flags: bit0 = run flood() on stalled particles (water.h:123-145), bit1 = full-grid seep pass
// (SoilMachine.cpp:300-301).  xy == NULL: spawn through the ctor's rand() as the reference does.
void smref_water_seq(int n, const float* xy, int flags, Stats* st) {
  memset(st, 0, sizeof(Stats));
  Layermap& map = *g.lmap;
  Vertexpool<Vertex>& vertexpool = *g.vp;
  double t0 = now();
  for (int i = 0; i < n; i++) {
    WaterParticle particle(map);
    if (xy) respawn(particle, map, xy[2 * i], xy[2 * i + 1]);
    while (true) {
      while (true) {
        if (!particle.move(map, vertexpool)) {
          if (particle.volume == 0.0) st->exit_oob++; else st->exit_stall++;
          break;
        }
        st->steps++;
        if (!particle.interact(map, vertexpool)) { st->exit_evap++; break; }
      }
      if (!(flags & 1)) break;
      if (!particle.flood(map, vertexpool)) break;
    }
  }
  st->seconds = now() - t0;
  if (flags & 2) WaterParticle::seep(map, vertexpool);
}
void smref_wind_seq(int n, const float* xy, int flags, Stats* st) {
  (void)flags;
  memset(st, 0, sizeof(Stats));
  Layermap& map = *g.lmap;
  Vertexpool<Vertex>& vertexpool = *g.vp;
  double t0 = now();
  for (int i = 0; i < n; i++) {
    WindParticle particle(map);
    if (xy) respawn(particle, map, xy[2 * i], xy[2 * i + 1]);
    while (true) {
      if (!particle.move(map, vertexpool)) { st->exit_oob++; break; }
      st->steps++;
      if (!particle.interact(map, vertexpool)) { st->exit_evap++; break; }
    }
  }
  st->seconds = now() - t0;
}

}  // extern "C"
