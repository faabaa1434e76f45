// oracle/sm_oracle.cpp -- plain CPU restatement of the reference's particle/terrain hot path.
//
// TEST INFRASTRUCTURE (see sm_oracle.h).  Written from the reference's behaviour, one function per
// reference function, each citing the lines it follows.  Data structures are deliberately the
// simplest possible (one std::vector of sections per cell, bottom -> top) so that this file can be
// audited against the reference without knowing anything about the CUDA layout.
//
// Arithmetic contract: types, promotions and association order are those of the reference
// expressions; GLM calls follow GLM 0.9.9 (see oracle/refharness/glm/glm.hpp); compile with
// -ffp-contract=off (the Makefile does).  Pinned bit for bit against oracle/_ref (the reference
// headers compiled verbatim) by tests/test_oracle_port.py.
#include "sm_oracle.h"
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <chrono>

namespace {

struct Section {        // struct sec, layermap.h:37-62 (links replaced by the vector order)
  int type;
  double size, floor, saturation;
};
typedef std::vector<Section> Column;

struct V2 { float x, y; };
struct V3 { float x, y, z; };

struct World {
  int dimx = 0, dimy = 0, SCALE = 80;
  std::vector<smo_soil> soils;       // soils[], surface.h:41-51
  std::vector<Column> col;           // dat[x*dim.y+y], layermap.h:131,151
  std::vector<float> wfreq, wtrack;  // WaterParticle::frequency / track, water.h:345-346
  std::vector<float> windfreq;       // WindParticle::frequency, wind.h:48
} W;

inline Column& at(int x, int y) { return W.col[(size_t)x * W.dimy + y]; }
const int AIR = 0;                   // soilmap["Air"], surface.h:53-57

// ---- Layermap queries ---------------------------------------------------------------------------
double height(int x, int y) {                       // Layermap::height(ivec2), layermap.h:422-425
  const Column& c = at(x, y);
  if (c.empty()) return 0.0;
  return c.back().floor + c.back().size;
}
int surface(int x, int y) {                         // Layermap::surface, layermap.h:417-420
  const Column& c = at(x, y);
  return c.empty() ? 0 : c.back().type;
}
double height_bilinear(V2 pos) {                    // Layermap::height(vec2), layermap.h:427-439
  double h = 0.0f;
  const float fx = std::floor(pos.x), fy = std::floor(pos.y);
  const int px = (int)fx, py = (int)fy;             // ivec2 p = floor(pos)
  const V2 w = {pos.x - fx, pos.y - fy};            // fract(pos)
  h += (1.0 - w.x) * (1.0 - w.y) * height(px, py);
  h += (1.0 - w.x) * w.y * height(px + 1, py);      // upstream's cross-wired weights, kept
  h += w.x * (1.0 - w.y) * height(px, py + 1);
  h += w.x * w.y * height(px + 1, py + 1);
  return h;
}
V3 cross(V3 x, V3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
V3 normal(int x, int y) {                           // Layermap::normal(ivec2), layermap.h:341-377
  V3 n = {0, 0, 0};
  const int S = W.SCALE;
  V3 p = {(float)x, (float)(S * height(x, y)), (float)y};
  int k = 0;
  if (x > 0 && y > 0) {
    V3 b = {(float)(x - 1), (float)(S * height(x - 1, y)), (float)y};
    V3 c = {(float)x, (float)(S * height(x, y - 1)), (float)(y - 1)};
    V3 r = cross(sub(c, p), sub(b, p));
    n.x += r.x; n.y += r.y; n.z += r.z; k++;
  }
  if (x > 0 && y < W.dimy - 1) {
    V3 b = {(float)(x - 1), (float)(S * height(x - 1, y)), (float)y};
    V3 c = {(float)x, (float)(S * height(x, y + 1)), (float)(y + 1)};
    V3 r = cross(sub(c, p), sub(b, p));
    n.x -= r.x; n.y -= r.y; n.z -= r.z; k++;
  }
  if (x < W.dimx - 1 && y > 0) {
    V3 b = {(float)(x + 1), (float)(S * height(x + 1, y)), (float)y};
    V3 c = {(float)x, (float)(S * height(x, y - 1)), (float)(y - 1)};
    V3 r = cross(sub(c, p), sub(b, p));
    n.x -= r.x; n.y -= r.y; n.z -= r.z; k++;
  }
  if (x < W.dimx - 1 && y < W.dimy - 1) {
    V3 b = {(float)(x + 1), (float)(S * height(x + 1, y)), (float)y};
    V3 c = {(float)x, (float)(S * height(x, y + 1)), (float)(y + 1)};
    V3 r = cross(sub(c, p), sub(b, p));
    n.x += r.x; n.y += r.y; n.z += r.z; k++;
  }
  const float fk = (float)k;
  n = {n.x / fk, n.y / fk, n.z / fk};
  const float inv = 1.0f / std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);   // glm::normalize
  return {n.x * inv, n.y * inv, n.z * inv};
}

// ---- Layermap mutators ----------------------------------------------------------------------------
void add(int x, int y, double size, int type, double saturation = 0.0) {   // Layermap::add, layermap.h:230-307
  Column& c = at(x, y);
  if (size <= 0) return;                                     // :237-240
  if (c.empty()) { c.push_back({type, size, 0.0, saturation}); return; }   // :243-246 (floor stays 0)
  if (c.back().type == type) { c.back().size += size; return; }             // :249-253
  if (c.back().type == AIR) {                                // :258-275 swap under the water section
    Section water = c.back();
    c.pop_back();
    add(x, y, size, type, saturation);
    add(x, y, water.size, AIR, water.saturation);
    return;
  }
  c.push_back({type, size, height(x, y), saturation});       // :302-305
}
double remove(int x, int y, double h) {                      // Layermap::remove, layermap.h:310-339
  Column& c = at(x, y);
  if (c.empty()) return 0.0;
  if (c.back().size <= 0.0) { c.pop_back(); return 0.0; }
  if (h <= 0.0) return 0.0;
  double diff = h - c.back().size;
  c.back().size -= h;
  if (diff >= 0.0) { c.pop_back(); return diff; }
  return 0.0;
}

// ---- mass budget -------------------------------------------------------------------------
// Six accumulators per particle-step, same definitions and same order of additions as the product's step
// (soilmachine_b200/csrc/sm_coop.cuh: eroded, deposited, cascade_net, discarded, clamped, wind_negative).  ACC
// points at the current step's six slots, or is null (hydrology, single-cell calls).
double* ACC = nullptr;

// ---- Particle::cascade, particle.h:24-101 ------------------------------------------------------------
void cascade(V2 pos, int transferloop) {
  const int ix = (int)std::round(pos.x), iy = (int)std::round(pos.y);
  static const int nx8[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
  static const int ny8[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
  struct Point { int x, y; double h; } sn[8];
  int num = 0;
  for (int k = 0; k < 8; k++) {
    const int nx = ix + nx8[k], ny = iy + ny8[k];
    if (nx >= W.dimx || ny >= W.dimy || nx < 0 || ny < 0) continue;
    sn[num++] = {nx, ny, height(nx, ny)};
  }
  // std::sort on <= 8 elements is libstdc++'s insertion sort = stable (particle.h:58-60)
  for (int i = 1; i < num; i++) {
    Point v = sn[i];
    int j = i;
    while (j > 0 && v.h > sn[j - 1].h) { sn[j] = sn[j - 1]; j--; }
    sn[j] = v;
  }
  for (int i = 0; i < num; i++) {
    const int nx = sn[i].x, ny = sn[i].y;
    float diff = (height(ix, iy) - height(nx, ny)) * (float)W.SCALE / 80.0f;   // :66
    if (diff == 0) continue;
    const int tx = (diff > 0) ? ix : nx, ty = (diff > 0) ? iy : ny;
    const int bx = (diff > 0) ? nx : ix, by = (diff > 0) ? ny : iy;
    const int type = surface(tx, ty);
    const smo_soil param = W.soils[type];
    float excess = std::fabs(diff) - param.maxdiff;
    if (excess <= 0) continue;
    float transfer = param.settling * excess / 2.0f;
    bool recascade = false;
    const double topsize = at(tx, ty).empty() ? 0.0 : at(tx, ty).back().size;
    if (transfer > topsize) transfer = topsize;              // :87-88 narrowing f64 -> f32
    const double ht0 = height(tx, ty), hb0 = height(bx, by);
    if (remove(tx, ty, transfer) != 0) recascade = true;
    add(bx, by, transfer, param.cascades);
    if (ACC) ACC[2] += (height(tx, ty) - ht0) + (height(bx, by) - hb0);
    if (recascade && transferloop > 0) cascade({(float)nx, (float)ny}, --transferloop);
  }
}

// ---- WaterParticle, water.h ------------------------------------------------------------------------------
struct Water {
  V2 pos, speed = {0, 0};
  double volume = 1.0, sediment = 0.0, evaprate = 0.001;
  int ix = 0, iy = 0;
  smo_soil param;
  int surf = 0, contains = 0;
  int spill = 3;                                             // water.h:33
};
void water_spawn(Water& p, float x, float y) {               // ctor, water.h:11-19
  p.pos = {x, y};
  p.ix = (int)std::round(x); p.iy = (int)std::round(y);
  p.surf = surface(p.ix, p.iy);
  p.param = W.soils[p.surf];
  p.contains = p.param.transports;
}
bool water_move(Water& p) {                                  // water.h:43-73
  p.ix = (int)std::round(p.pos.x); p.iy = (int)std::round(p.pos.y);
  const V3 n = normal(p.ix, p.iy);
  p.surf = surface(p.ix, p.iy);
  p.param = W.soils[p.surf];
  p.evaprate = 0.01;
  const int ind = p.iy * W.dimx + p.ix;
  W.wtrack[ind] += p.volume;                                 // updatefrequency, :348-351
  p.param.friction = p.param.friction * (1.0f - W.wfreq[ind]);
  p.evaprate = p.evaprate * (1.0f - 0.2f * W.wfreq[ind]);
  {
    const V2 v = {n.x * p.param.friction, n.z * p.param.friction};
    if (std::sqrt(v.x * v.x + v.y * v.y) < 1E-5) return false;
  }
  const float f = p.param.friction;                          // mix(vec2(n.x,n.z), speed, friction)
  V2 s = {n.x * (1.0f - f) + p.speed.x * f, n.z * (1.0f - f) + p.speed.y * f};
  const float inv = 1.0f / std::sqrt(s.x * s.x + s.y * s.y);
  const float r2 = std::sqrt(2.0f);
  p.speed = {r2 * (s.x * inv), r2 * (s.y * inv)};
  p.pos.x += p.speed.x; p.pos.y += p.speed.y;
  if (!(p.pos.x >= 0.0f && p.pos.y >= 0.0f) ||
      !(p.pos.x < (float)W.dimx - 1.0f && p.pos.y < (float)W.dimy - 1.0f)) {
    if (ACC) ACC[3] += p.sediment * p.volume;
    p.volume = 0.0;
    return false;
  }
  return true;
}
bool water_interact(Water& p) {                              // water.h:75-121
  double c_eq = p.param.solubility * (height(p.ix, p.iy) - height_bilinear(p.pos)) * (double)W.SCALE / 80.0;
  if (c_eq < 0.0) c_eq = 0.0;
  if (c_eq > 1.0) c_eq = 1.0;
  const int ind = p.iy * W.dimx + p.ix;
  if ((double)(W.soils[p.contains].erosionrate) < W.wfreq[ind]) p.contains = W.soils[p.contains].erodes;
  const double cdiff = c_eq - p.sediment;
  if (cdiff > 0) {
    p.sediment += p.param.equrate * cdiff;
    p.contains = W.soils[surface(p.ix, p.iy)].transports;
    const double h0 = height(p.ix, p.iy);
    double diff = remove(p.ix, p.iy, p.param.equrate * cdiff * p.volume);
    while (std::fabs(diff) > 1E-8) diff = remove(p.ix, p.iy, diff);
    if (ACC) ACC[0] += h0 - height(p.ix, p.iy);
  } else if (cdiff < 0) {
    p.sediment += W.soils[p.contains].equrate * cdiff;
    const double h0 = height(p.ix, p.iy);
    add(p.ix, p.iy, -W.soils[p.contains].equrate * cdiff * p.volume, p.contains);
    if (ACC) ACC[1] += height(p.ix, p.iy) - h0;
  }
  cascade(p.pos, 0);
  p.sediment /= (1.0 - p.evaprate);
  const double over = p.sediment - 1.0;
  if (p.sediment > 1.0) p.sediment = 1.0;
  p.volume *= (1.0 - p.evaprate);
  if (ACC) {
    if (over > 0.0) ACC[4] += over * p.volume;
    if (!(p.volume > 0.01)) ACC[3] += p.sediment * p.volume;
  }
  return p.volume > 0.01;
}

// ---- pooling hydrology, water.h:123-343 ---------------------------------
const double volumeFactor = 0.015;                           // water.h:368
smo_hydro H;                                                 // counters of the current hydrology call
void water_cascade(int ix, int iy, int spill);

void water_seep(int x, int y) {                              // WaterParticle::seep(vec2,...), water.h:285-333
  Column& c = at(x, y);
  if (c.empty()) return;
  // `top` walks down the column; sections are named by their index from the bottom, which a pop of the
  // column's top (map.remove acts on dat[], not on `top`) leaves valid for everything underneath.
  for (int t = (int)c.size() - 1; t >= 1; t--) {
    const smo_soil param = W.soils[c[t].type], nparam = W.soils[c[t - 1].type];
    const double vol = c[t].size * c[t].saturation * param.porosity;
    const double nevol = c[t - 1].size * (1.0 - c[t - 1].saturation) * nparam.porosity;
    double seepage = 1.0;
    const double transfer = (vol < nevol) ? vol : nevol;
    if (transfer < 1E-6) seepage = 1.0;
    if (transfer > 0) {
      if (c[t].type == AIR) remove(x, y, seepage * transfer);
      else c[t].saturation -= (seepage * transfer) / (c[t].size * param.porosity);
      c[t - 1].saturation += (seepage * transfer) / (c[t - 1].size * nparam.porosity);
    }
  }
}

bool water_flood(Water& p) {                                 // WaterParticle::flood, water.h:123-145
  if (p.volume < 0.01 || p.spill-- <= 0) return false;
  H.floods++;
  p.ix = (int)p.pos.x; p.iy = (int)p.pos.y;                  // ipos = pos truncates (:128)
  add(p.ix, p.iy, p.sediment * W.soils[p.contains].equrate, p.contains);
  cascade(p.pos, 0);
  add(p.ix, p.iy, p.volume * volumeFactor, AIR);
  water_seep(p.ix, p.iy);
  water_cascade(p.ix, p.iy, p.spill);
  return false;
}

void water_to_completion(Water& p, int64_t* steps) {         // SoilMachine.cpp:292-296 / water.h:252-256
  for (;;) {
    while (water_move(p)) {
      ++*steps;
      if (!water_interact(p)) break;
    }
    if (!water_flood(p)) break;
  }
}

void water_cascade(int ix, int iy, int spill) {              // WaterParticle::cascade, water.h:151-283
  static const int nx8[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
  static const int ny8[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
  struct Point { int x, y; double h; } sn[8];
  int num = 0;
  for (int k = 0; k < 8; k++) {
    const int nx = ix + nx8[k], ny = iy + ny8[k];
    if (nx >= W.dimx || ny >= W.dimy || nx < 0 || ny < 0) continue;
    sn[num++] = {nx, ny, height(nx, ny)};
  }
  for (int i = 1; i < num; i++) {                            // std::sort on <= 8 elements, highest first
    Point v = sn[i];
    int j = i;
    while (j > 0 && v.h > sn[j - 1].h) { sn[j] = sn[j - 1]; j--; }
    sn[j] = v;
  }
  for (int i = 0; i < num; i++) {
    const int nx = sn[i].x, ny = sn[i].y;
    const Column& A = at(ix, iy);
    const Column& B = at(nx, ny);
    double whA = 0, whB = 0, fA = 0.0, fB = 0.0;             // water table = top of the column (:183-203)
    if (!A.empty()) { whA = A.back().size; fA = A.back().floor; }
    if (!B.empty()) { whB = B.back().size; fB = B.back().floor; }
    const double diff = (fA + whA - fB - whB) * (double)W.SCALE / 80.0;
    if (diff == 0) continue;
    const int tx = (diff > 0) ? ix : nx, ty = (diff > 0) ? iy : ny;
    const int bx = (diff > 0) ? nx : ix, by = (diff > 0) ? ny : iy;
    const Column& top = at(tx, ty);
    if (top.empty() || top.back().type != AIR) continue;     // only water moves (:218-219)
    double transfer = std::fabs(diff) / 2.0;
    const double wh = top.back().size;
    transfer = (wh < transfer) ? wh : transfer;
    if (transfer <= 0) continue;
    bool recascade = false;
    if (transfer == wh) {                                    // the whole water section leaves as a particle
      remove(tx, ty, transfer);
      Water q;
      // the ctor draws a random position only to read `contains` there (water.h:13-17); that value is
      // overwritten by the first erosion before it can reach the map (deposits need sediment > 0)
      water_spawn(q, (float)tx, (float)ty);
      const V2 d = {(float)bx - (float)tx, (float)by - (float)ty};
      const float inv = 1.0f / std::sqrt(d.x * d.x + d.y * d.y);
      const float r2 = std::sqrt(2.0f);
      q.speed = {r2 * (d.x * inv), r2 * (d.y * inv)};
      q.spill = spill;
      q.volume = transfer / volumeFactor;
      H.nested++;
      water_to_completion(q, &H.nested_steps);
    } else {
      if (remove(tx, ty, transfer) != 0) recascade = true;
      if (transfer > 0) recascade = true;
      add(bx, by, transfer, AIR);
      at(bx, by).back().saturation = 1.0f;
      H.transfers++;
    }
    if (recascade && spill > 0) water_cascade(nx, ny, --spill);
  }
}

// ---- WindParticle, wind.h ----------------------------------------------------------------------------------
struct Wind {
  V2 pos;
  V3 speed = {-2, 0, 1};
  double sediment = 0.0, height = 0.0, sheight = 0.0;
  int ix = 0, iy = 0, surf = 0, contains = 0;
  smo_soil param;
};
void wind_spawn(Wind& p, float x, float y) {                 // ctor, wind.h:13-22
  p.pos = {x, y};
  p.ix = (int)std::round(x); p.iy = (int)std::round(y);
  p.surf = surface(p.ix, p.iy);
  p.param = W.soils[p.surf];
  p.contains = p.param.transports;
}
// Prevailing wind.  Upstream: the constant (-2, 0, 1) (wind.h:29).  Extension (off unless a field is attached): the
// nearest cell of a lattice velocity field, lattice <-> world as SoilMachine.cpp:234-239 maps them, velocity / 0.05,
// components clamped to [-2, 2]; the product's rule is soilmachine_b200/csrc/sm_coop.cuh wind_field_pspeed.
std::vector<float> FIELD; int FNX = 0, FNY = 0, FNZ = 0;
V3 pspeed_at(const V2& pos, double height) {
  if (FIELD.empty()) return {-2, 0, 1};
  const float sx = (float)W.dimx / (float)FNX, sy = (float)W.SCALE / 32.0f, sz = (float)W.dimy / (float)FNZ;
  int lx = (int)(pos.x / sx), lz = (int)(pos.y / sz);
  int ly = (int)((float)(height * 80.0 / (double)W.SCALE) * (float)W.SCALE / sy);
  lx = std::min(std::max(lx, 0), FNX - 1); ly = std::min(std::max(ly, 0), FNY - 1); lz = std::min(std::max(lz, 0), FNZ - 1);
  const float* v = &FIELD[(((size_t)lx * FNY + ly) * FNZ + lz) * 4];
  float c[3];
  for (int k = 0; k < 3; k++) { c[k] = v[k] / 0.05f; c[k] = c[k] < -2.0f ? -2.0f : (c[k] > 2.0f ? 2.0f : c[k]); }
  return {c[0], c[1], c[2]};
}
V3 mixd(V3 x, V3 y, double a) {                              // glm::mix with a double weight
  return {(float)((double)x.x * (1.0 - a) + (double)y.x * a), (float)((double)x.y * (1.0 - a) + (double)y.y * a),
          (float)((double)x.z * (1.0 - a) + (double)y.z * a)};
}
bool wind_move(Wind& p) {                                    // wind.h:54-92
  if (W.soils[p.contains].suspension == 0.0) return false;
  p.ix = (int)std::round(p.pos.x); p.iy = (int)std::round(p.pos.y);
  const V3 n = normal(p.ix, p.iy);
  p.surf = surface(p.ix, p.iy);
  p.param = W.soils[p.surf];
  const int ind = p.iy * W.dimx + p.ix;
  W.windfreq[ind] = 0.5 * W.windfreq[ind] + 0.5f;            // updatefrequency, :49-52
  p.sheight = height(p.ix, p.iy) * (float)W.SCALE / 80.0f;
  if (p.height < p.sheight) p.height = p.sheight;
  if (p.height > p.sheight) p.speed.y -= 0.25;
  else p.speed = mixd(p.speed, cross(cross(p.speed, n), n), 0.8);
  p.speed = mixd(p.speed, pspeed_at(p.pos, p.height), 0.2);
  p.pos.x += p.speed.x; p.pos.y += p.speed.z;
  p.height += p.speed.y;
  if (!(p.pos.x >= 0.0f && p.pos.y >= 0.0f) || !((int)p.pos.x < W.dimx - 1 && (int)p.pos.y < W.dimy - 1) ||
      std::sqrt(p.speed.x * p.speed.x + p.speed.y * p.speed.y + p.speed.z * p.speed.z) < 0.01) {
    if (ACC) ACC[3] += p.sediment;
    return false;
  }
  return true;
}
bool wind_interact(Wind& p) {                                // wind.h:94-136
  const int nx = (int)std::round(p.pos.x), ny = (int)std::round(p.pos.y);
  if (p.height <= height_bilinear(p.pos) * (float)W.SCALE / 80.0f) {
    if (p.param.transports == p.contains) {
      const float len = std::sqrt(p.speed.x * p.speed.x + p.speed.y * p.speed.y + p.speed.z * p.speed.z);
      double force = len * (height(nx, ny) - p.height) * (float)W.SCALE / 80.0f * (1.0f - p.sediment);
      const double h0 = height(p.ix, p.iy);
      double diff = remove(p.ix, p.iy, p.param.suspension * force);
      if (ACC) { ACC[0] += h0 - height(p.ix, p.iy); if (p.param.suspension * force < 0.0) ACC[5] += p.param.suspension * force; }
      p.sediment += (p.param.suspension * force - diff);
      cascade({(float)p.ix, (float)p.iy}, 1);
    }
  } else if (p.param.suspension > 0.0) {
    p.sediment -= W.soils[p.contains].suspension * p.sediment;
    double h0 = height(nx, ny);
    add(nx, ny, 0.5f * W.soils[p.contains].suspension * p.sediment, p.contains);
    if (ACC) ACC[1] += height(nx, ny) - h0;
    h0 = height(p.ix, p.iy);
    add(p.ix, p.iy, 0.5f * W.soils[p.contains].suspension * p.sediment, p.contains);
    if (ACC) ACC[1] += height(p.ix, p.iy) - h0;
    cascade({(float)p.ix, (float)p.iy}, 1);
    cascade({(float)nx, (float)ny}, 1);
  }
  return true;
}

std::vector<Water> WP; std::vector<int> Wlive;
std::vector<Wind> DP; std::vector<int> Dlive;
std::vector<double> BUD;          // 6 accumulators per particle of the current lockstep batch
// run one particle-step with the budget attached: the step's six sums are added to the particle's totals
template <class F> inline void with_budget(int i, F f) {
  double acc[6] = {0, 0, 0, 0, 0, 0};
  ACC = acc;
  f();
  ACC = nullptr;
  for (int k = 0; k < 6; k++) BUD[(size_t)i * 6 + k] += acc[k];
}
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

extern "C" {

void smo_init(int dimx, int dimy, int scale, int nsoils, const smo_soil* soils) {
  W = World();
  W.dimx = dimx; W.dimy = dimy; W.SCALE = scale;
  W.soils.assign(soils, soils + nsoils);
  W.col.assign((size_t)dimx * dimy, Column());
  W.wfreq.assign((size_t)dimx * dimy, 0.0f); W.wtrack = W.wfreq; W.windfreq = W.wfreq;
  WP.clear(); DP.clear(); Wlive.clear(); Dlive.clear();
}
void smo_set_columns(const int64_t* off, const int32_t* type, const double* size, const double* sat) {
  for (int x = 0; x < W.dimx; x++) for (int y = 0; y < W.dimy; y++) {
    const size_t c = (size_t)x * W.dimy + y;
    W.col[c].clear();
    for (int64_t k = off[c]; k < off[c + 1]; k++) add(x, y, size[k], type[k], sat ? sat[k] : 0.0);
  }
}
int64_t smo_nsections(void) { int64_t n = 0; for (auto& c : W.col) n += (int64_t)c.size(); return n; }
void smo_get_columns(int64_t* off, int32_t* type, double* size, double* floor_, double* sat) {
  int64_t n = 0;
  for (size_t c = 0; c < W.col.size(); c++) {
    off[c] = n;
    for (const Section& s : W.col[c]) { type[n] = s.type; size[n] = s.size; floor_[n] = s.floor; sat[n] = s.saturation; n++; }
  }
  off[W.col.size()] = n;
}
void smo_heights(double* out) {
  for (int x = 0; x < W.dimx; x++) for (int y = 0; y < W.dimy; y++) out[(size_t)x * W.dimy + y] = height(x, y);
}
void smo_get_frequency(float* a, float* b, float* c) {
  const size_t n = W.wfreq.size() * 4;
  if (a) memcpy(a, W.wfreq.data(), n); if (b) memcpy(b, W.wtrack.data(), n); if (c) memcpy(c, W.windfreq.data(), n);
}
void smo_set_frequency(const float* a, const float* b, const float* c) {
  const size_t n = W.wfreq.size() * 4;
  if (a) memcpy(W.wfreq.data(), a, n); if (b) memcpy(W.wtrack.data(), b, n); if (c) memcpy(W.windfreq.data(), c, n);
}
void smo_frequency_update(void) {                            // mapfrequency + resetfrequency, water.h:353-365
  const float lrate = 0.01f, K = 50.0f;
  for (size_t i = 0; i < W.wfreq.size(); i++)
    W.wfreq[i] = (1.0f - lrate) * W.wfreq[i] + lrate * K * W.wtrack[i] / (1.0f + K * W.wtrack[i]);
  for (size_t i = 0; i < W.wtrack.size(); i++) W.wtrack[i] = 0.0f;
}
double smo_height_i(int x, int y) { return height(x, y); }
double smo_height_f(float x, float y) { return height_bilinear({x, y}); }
int smo_surface(int x, int y) { return surface(x, y); }
void smo_normal(int x, int y, float* o) { V3 n = normal(x, y); o[0] = n.x; o[1] = n.y; o[2] = n.z; }
void smo_add(int x, int y, double size, int type) { add(x, y, size, type); }
double smo_remove(int x, int y, double h) { return remove(x, y, h); }
void smo_cascade(float x, float y, int loop) { cascade({x, y}, loop); }

void smo_water_begin(int n, const float* xy) {
  WP.assign(n, Water()); Wlive.clear(); BUD.assign((size_t)n * 6, 0.0);
  for (int i = 0; i < n; i++) { water_spawn(WP[i], xy[2 * i], xy[2 * i + 1]); Wlive.push_back(i); }
}
int smo_water_sweep(smo_stats* st) {
  std::vector<int> next;
  for (int i : Wlive) {
    Water& p = WP[i];
    bool moved = false, lives = false;
    with_budget(i, [&]() { moved = water_move(p); if (moved) lives = water_interact(p); });
    if (!moved) { if (p.volume == 0.0) st->exit_oob++; else st->exit_stall++; continue; }
    st->steps++;
    if (!lives) { st->exit_evap++; continue; }
    next.push_back(i);
  }
  Wlive.swap(next); st->sweeps++;
  return (int)Wlive.size();
}
void smo_set_wind_field(const float* v4, int nx, int ny, int nz) {
  if (!v4) { FIELD.clear(); return; }
  FIELD.assign(v4, v4 + (size_t)nx * ny * nz * 4); FNX = nx; FNY = ny; FNZ = nz;
}
// mass budget of the current lockstep batch: per-particle accumulators (n x 6) and their sums in particle order
int64_t smo_budget(double* per_particle, double* sums6) {
  const size_t n = BUD.size() / 6;
  if (per_particle) memcpy(per_particle, BUD.data(), BUD.size() * sizeof(double));
  if (sums6) {
    for (int k = 0; k < 6; k++) sums6[k] = 0.0;
    for (size_t i = 0; i < n; i++) for (int k = 0; k < 6; k++) sums6[k] += BUD[i * 6 + k];
  }
  return (int64_t)n;
}
void smo_water_state(float* pos, float* speed, double* vol, double* sed, int32_t* cont, int32_t* alive) {
  for (size_t i = 0; i < WP.size(); i++) {
    pos[2 * i] = WP[i].pos.x; pos[2 * i + 1] = WP[i].pos.y; speed[2 * i] = WP[i].speed.x; speed[2 * i + 1] = WP[i].speed.y;
    vol[i] = WP[i].volume; sed[i] = WP[i].sediment; cont[i] = WP[i].contains; alive[i] = 0;
  }
  for (int i : Wlive) alive[i] = 1;
}
void smo_water_run(int n, const float* xy, int max_sweeps, smo_stats* st) {
  memset(st, 0, sizeof(*st));
  smo_water_begin(n, xy);
  const double t0 = now();
  while (!Wlive.empty() && (max_sweeps <= 0 || st->sweeps < max_sweeps)) smo_water_sweep(st);
  st->seconds = now() - t0;
}
// floods of the finished lockstep batch, ascending particle index (see oracle/refharness: smref_water_flood)
void smo_water_flood(smo_hydro* out) {
  H = smo_hydro();
  std::vector<char> live(WP.size(), 0);
  for (int i : Wlive) live[i] = 1;
  for (size_t i = 0; i < WP.size(); i++) if (!live[i]) water_flood(WP[i]);
  if (out) *out = H;
}
// WaterParticle::seep(map, vertexpool), water.h:335-343 / SoilMachine.cpp:300-301
void smo_seep(smo_hydro* out) {
  H = smo_hydro();
  for (int x = 0; x < W.dimx; x++) for (int y = 0; y < W.dimy; y++) {
    water_seep(x, y);
    water_cascade(x, y, 3);
    H.cells++;
  }
  if (out) *out = H;
}
void smo_wind_begin(int n, const float* xy) {
  DP.assign(n, Wind()); Dlive.clear(); BUD.assign((size_t)n * 6, 0.0);
  for (int i = 0; i < n; i++) { wind_spawn(DP[i], xy[2 * i], xy[2 * i + 1]); Dlive.push_back(i); }
}
int smo_wind_sweep(smo_stats* st) {
  std::vector<int> next;
  for (int i : Dlive) {
    Wind& p = DP[i];
    bool moved = false, lives = false;
    with_budget(i, [&]() { moved = wind_move(p); if (moved) lives = wind_interact(p); });
    if (!moved) { st->exit_oob++; continue; }
    st->steps++;
    if (!lives) { st->exit_evap++; continue; }
    next.push_back(i);
  }
  Dlive.swap(next); st->sweeps++;
  return (int)Dlive.size();
}
void smo_wind_state(float* pos, float* speed3, double* h, double* sed, int32_t* cont, int32_t* alive) {
  for (size_t i = 0; i < DP.size(); i++) {
    pos[2 * i] = DP[i].pos.x; pos[2 * i + 1] = DP[i].pos.y;
    speed3[3 * i] = DP[i].speed.x; speed3[3 * i + 1] = DP[i].speed.y; speed3[3 * i + 2] = DP[i].speed.z;
    h[i] = DP[i].height; sed[i] = DP[i].sediment; cont[i] = DP[i].contains; alive[i] = 0;
  }
  for (int i : Dlive) alive[i] = 1;
}
void smo_wind_run(int n, const float* xy, int max_sweeps, smo_stats* st) {
  memset(st, 0, sizeof(*st));
  smo_wind_begin(n, xy);
  const double t0 = now();
  while (!Dlive.empty() && (max_sweeps <= 0 || st->sweeps < max_sweeps)) smo_wind_sweep(st);
  st->seconds = now() - t0;
}
void smo_water_seq(int n, const float* xy, smo_stats* st) {  // SoilMachine.cpp:288-298 without flood
  memset(st, 0, sizeof(*st));
  const double t0 = now();
  for (int i = 0; i < n; i++) {
    Water p;
    water_spawn(p, xy[2 * i], xy[2 * i + 1]);
    for (;;) {
      if (!water_move(p)) { if (p.volume == 0.0) st->exit_oob++; else st->exit_stall++; break; }
      st->steps++;
      if (!water_interact(p)) { st->exit_evap++; break; }
    }
  }
  st->seconds = now() - t0;
}
// the full water part of the reference frame: flags bit0 = flood (SoilMachine.cpp:292-296), bit1 = seep pass (:300-301)
void smo_water_seq_full(int n, const float* xy, int flags, smo_stats* st, smo_hydro* out) {
  memset(st, 0, sizeof(*st));
  H = smo_hydro();
  const double t0 = now();
  for (int i = 0; i < n; i++) {
    Water p;
    water_spawn(p, xy[2 * i], xy[2 * i + 1]);
    for (;;) {
      for (;;) {
        if (!water_move(p)) { if (p.volume == 0.0) st->exit_oob++; else st->exit_stall++; break; }
        st->steps++;
        if (!water_interact(p)) { st->exit_evap++; break; }
      }
      if (!(flags & 1)) break;
      if (!water_flood(p)) break;
    }
  }
  st->seconds = now() - t0;
  if (flags & 2) for (int x = 0; x < W.dimx; x++) for (int y = 0; y < W.dimy; y++) { water_seep(x, y); water_cascade(x, y, 3); H.cells++; }
  if (out) *out = H;
}
void smo_wind_seq(int n, const float* xy, smo_stats* st) {   // SoilMachine.cpp:304-307
  memset(st, 0, sizeof(*st));
  const double t0 = now();
  for (int i = 0; i < n; i++) {
    Wind p;
    wind_spawn(p, xy[2 * i], xy[2 * i + 1]);
    for (;;) {
      if (!wind_move(p)) { st->exit_oob++; break; }
      st->steps++;
      if (!wind_interact(p)) { st->exit_evap++; break; }
    }
  }
  st->seconds = now() - t0;
}

}  // extern "C"
