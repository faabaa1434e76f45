// sm_hydro.cuh -- pooling hydrology: flood, water-table cascade, seep.
//
// Reference map:  WaterParticle::flood                 source/particle/water.h:123-145
//                 WaterParticle::cascade (water table)  source/particle/water.h:151-283
//                 WaterParticle::seep(vec2) / seep(map) source/particle/water.h:285-343
//                 frame loop                            SoilMachine.cpp:292-301
//
// Written against the same accessor contract as sm_core.cuh (plus a.wet_mark(x,y)), so the code below
// is compiled into the device executor (sm_engine.cu: k_hydro_*) and, for the CPU test-suite, into
// tests/hostsim.
//
// Upstream is recursive here: flood -> water cascade -> nested WaterParticle run to completion -> its
// flood -> water cascade -> ... and cascade -> cascade(neighbour, --spill).  Two facts flatten that into a
// loop over a small explicit stack of cascade frames:
//   * flood() does nothing after its cascade call (water.h:139-143), and a nested particle does nothing
//     after its flood (water.h:252-256): both are tail positions, so only cascade frames need to be kept;
//   * every child frame starts with a spill budget at least one below its parent's current budget
//     (recursion: --spill, water.h:277-278; nested particle: flood's spill-- , water.h:125), and budgets
//     start at 3, so at most four frames are ever open.
#pragma once
#include "sm_core.cuh"

#define SM_VOLUME_FACTOR 0.015   // default of WaterParticle::volumeFactor (water.h:368, a mutable static upstream: a.volume_factor())
#define SM_MINVOL 0.01           // WaterParticle::minvol, water.h:31
#define SM_WSTACK 6              // open water-cascade frames (4 needed, see above)

struct HydroCount {              // counters of one hydrology call
  unsigned long long floods;        // flood() calls that passed the volume/spill guard (nested ones included)
  unsigned long long nested;        // particles spawned by the water-table cascade (water.h:243-256)
  unsigned long long nested_steps;  // their particle-steps
  unsigned long long transfers;     // partial water-table transfers (water.h:260-272)
  unsigned long long cells;         // seep pass: cells visited
  unsigned long long overflow;      // frame-stack overflow (cannot happen; checked by the callers)
};

// One open WaterParticle::cascade call: centre cell, the in-bounds neighbours in visiting order (nibble i of
// `order` = neighbour index k with rank i), the next rank to visit and the remaining spill budget.
struct WFrame {
  int cx, cy;
  unsigned int order;
  int num, i, spill;
};

// ------------------------------------------------------------------------------------------------
// WaterParticle::seep(vec2 pos, ...), water.h:285-333: pass water from each section to the one below
// ------------------------------------------------------------------------------------------------
// Sections are named by handles: SM_H_TOP = the cell's top record, anything else = a pool slot.  A
// pop moves the section under the top from its pool slot into the top record (col_pop), so the
// handle of `prev` is re-pointed when that happens; upstream's pointers stay valid across the pop for
// the same reason (the section is not moved there).
#define SM_H_TOP 0xFFFFFFFEu
template <class A> SM_HD_NOINLINE void hydro_seep_cell(A& a, int x, int y) {
  Sec32* const r = a.rec(x, y);
  if (r->type == SM_EMPTY) return;                                  // :290-292
  a.focus(x, y);
  Sec32 cur = *r;
  uint32_t hcur = SM_H_TOP;
  bool touched = false;
  SM_UNROLL1
  while (cur.below != SM_NIL) {                                     // :294  top->prev != NULL
    uint32_t hprev = cur.below;
    Sec32 prev = a.pool_load(hprev);
    const SoilDev param = a.soil(cur.type), nparam = a.soil(prev.type);   // :298-299
    const double vol = cur.size * cur.saturation * param.porosity;         // :302
    const double nevol = prev.size * (1.0 - prev.saturation) * nparam.porosity;   // :308
    const double transfer = (vol < nevol) ? vol : nevol;            // :313 (seepage stays 1.0, :310,314)
    if (transfer > 0) {                                             // :316
      if (cur.type == SM_AIR) {                                     // :319-320 map.remove acts on dat[]
        const uint32_t below_before = r->below;
        col_remove(a, *r, transfer);
        const bool popped = (r->type == SM_EMPTY) || (r->below != below_before);
        if (popped && hprev == below_before) hprev = SM_H_TOP;      // prev has become the top record
      } else {                                                      // :321-322
        cur.saturation -= transfer / (cur.size * param.porosity);
        if (hcur == SM_H_TOP) r->saturation = cur.saturation; else a.pool_store(hcur, cur);
      }
      prev.saturation += transfer / (prev.size * nparam.porosity);   // :324
      if (hprev == SM_H_TOP) r->saturation = prev.saturation; else a.pool_store(hprev, prev);
      touched = true;
    }
    hcur = hprev;                                                   // :328
    cur = prev;
  }
  if (touched) { a.wet_mark(x, y); a.dirty_rec(r, x, y); }
}

// ------------------------------------------------------------------------------------------------
// WaterParticle::cascade, water.h:151-283, as a frame machine
// ------------------------------------------------------------------------------------------------
// gather the in-bounds neighbours and their visiting order (water.h:155-183): highest first, stable
template <class A> SM_HD_NOINLINE void hydro_open(A& a, WFrame& f, int cx, int cy, int spill) {
  const int dimx = a.dimx(), dimy = a.dimy();
  double h[8];
  int num = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int kk = k + (k >= 4 ? 1 : 0);
    const int nx = cx + kk / 3 - 1, ny = cy + kk % 3 - 1;
    const bool in = !(nx >= dimx || ny >= dimy || nx < 0 || ny < 0);   // :171-172
    h[k] = -1.0e300;
    if (in) { h[k] = map_height(a, nx, ny); num++; }
  }
  int rank[8];
#pragma unroll
  for (int k = 0; k < 8; k++) rank[k] = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
#pragma unroll
    for (int j = k + 1; j < 8; j++) {
      const bool j_first = h[j] > h[k];
      rank[k] += j_first ? 1 : 0;
      rank[j] += j_first ? 0 : 1;
    }
  }
  unsigned int order = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) order |= (unsigned int)k << (4 * rank[k]);
  f.cx = cx; f.cy = cy; f.order = order; f.num = num; f.i = 0; f.spill = spill;
}

template <class A> SM_HD void hydro_push(A& a, WFrame* st, int& sp, int cx, int cy, int spill, HydroCount& hc) {
  if (sp >= SM_WSTACK) { hc.overflow++; return; }
  hydro_open(a, st[sp], cx, cy, spill);
  sp++;
}

// WaterParticle::flood, water.h:123-145.  Its trailing cascade call becomes a pushed frame.
template <class A> SM_HD_NOINLINE void hydro_flood(A& a, const WaterP& p, int spill, WFrame* st, int& sp, HydroCount& hc) {
  if (p.volume < SM_MINVOL || spill-- <= 0) return;                 // :125-126
  hc.floods++;
  const int ix = (int)p.px, iy = (int)p.py;                         // :128 ipos = pos truncates
  Sec32* const r = a.rec(ix, iy);
  a.focus(ix, iy);
  col_add(a, *r, p.sediment * a.soil(p.contains).equrate, p.contains);   // :133
  a.dirty_rec(r, ix, iy);
  Cascade<0, A>::run(a, (int)roundf(p.px), (int)roundf(p.py), 0);   // :134
  a.focus(ix, iy);
  col_add(a, *r, p.volume * a.volume_factor(), SM_AIR);            // :138
  a.dirty_rec(r, ix, iy);
  hydro_seep_cell(a, ix, iy);                                       // :139
  hydro_push(a, st, sp, ix, iy, spill, hc);                         // :140
}

// the nested particle's move() && interact() (water.h:252-253), out of line: it is the rare branch of the
// frame loop and would otherwise triple the loop's code size
template <class A> SM_HD_NOINLINE int hydro_nested_step(A& a, WaterP& q) { return water_step(a, q); }

// run every open frame to its end
template <class A> SM_HD_NOINLINE void hydro_drain(A& a, WFrame* st, int& sp, HydroCount& hc) {
  const int SCALE = a.scale();
  SM_UNROLL1
  while (sp > 0) {
    WFrame& f = st[sp - 1];
    if (f.i >= f.num) { sp--; continue; }
    const int i = f.i++;
    const int k = (int)((f.order >> (4 * i)) & 7u);
    const int kk = k + (k >= 4 ? 1 : 0);
    const int cx = f.cx, cy = f.cy;
    const int nx = cx + kk / 3 - 1, ny = cy + kk % 3 - 1;
    Sec32* const pa = a.rec(cx, cy);                                // secA = map.top(ipos), :185-186
    Sec32* const pb = a.rec(nx, ny);
    double whA = 0, whB = 0, fA = 0.0, fB = 0.0;                    // :189-208
    if (pa->type != SM_EMPTY) { whA = pa->size; fA = pa->floor; }
    if (pb->type != SM_EMPTY) { whB = pb->size; fB = pb->floor; }
    // diff = num / 80.0 (:211) has num's sign and is zero only when num is zero or the quotient underflows;
    // both early exits below are side-effect free, so the IEEE division is only paid where water can move
    const double num = (fA + whA - fB - whB) * (double)SCALE;
    if (num == 0) continue;
    Sec32* const top = (num > 0) ? pa : pb;                         // :216-220
    Sec32* const bot = (num > 0) ? pb : pa;
    const int tx = (num > 0) ? cx : nx, ty = (num > 0) ? cy : ny;
    const int bx = (num > 0) ? nx : cx, by = (num > 0) ? ny : cy;
    if (top->type != SM_AIR) continue;                              // :223-224 only water moves
    const double diff = num / 80.0;
    if (diff == 0) continue;                                        // :212-213
    double transfer = fabs(diff) / 2.0;                             // :227
    const double wh = top->size;                                    // :230
    transfer = (wh < transfer) ? wh : transfer;
    if (transfer <= 0) continue;                                    // :233-234
    if (transfer == wh) {                                           // :240-258 all of it leaves as a particle
      a.focus(tx, ty);
      col_remove(a, *top, transfer);
      a.dirty_rec(top, tx, ty);
      WaterP q;
      q.px = (float)tx; q.py = (float)ty;
      {
        const float dx = (float)bx - (float)tx, dy = (float)by - (float)ty;   // :246
        const float inv = 1.0f / sqrtf(dx * dx + dy * dy);
        q.sx = SM_SQRT2F * (dx * inv);
        q.sy = SM_SQRT2F * (dy * inv);
      }
      q.volume = transfer / a.volume_factor();                     // :250
      q.sediment = 0.0;
      // the ctor reads `contains` at a rand() position (water.h:13-17); the value cannot reach the map:
      // deposits need sediment > 0, which only an erosion produces, and every erosion re-derives contains
      q.contains = a.soil(rec_surface(*top)).transports;
      const int qspill = f.spill;                                   // :249
      hc.nested++;
      SM_UNROLL1
      for (;;) {                                                    // :252-253
        const int rc = hydro_nested_step(a, q);
        if (rc == SM_ALIVE || rc == SM_EXIT_EVAP) hc.nested_steps++;
        if (rc != SM_ALIVE) break;
      }
      hydro_flood(a, q, qspill, st, sp, hc);                        // :254
    } else {                                                        // :260-272
      a.focus(tx, ty);
      col_remove(a, *top, transfer);
      a.dirty_rec(top, tx, ty);
      a.focus(bx, by);
      col_add(a, *bot, transfer, SM_AIR);
      if (bot->type != SM_EMPTY) bot->saturation = 1.0;             // map.top(bpos)->saturation = 1.0f
      a.wet_mark(bx, by);
      a.dirty_rec(bot, bx, by);
      hc.transfers++;
      if (f.spill > 0) {                                            // :277-278 cascade(npos, --spill)
        --f.spill;
        hydro_push(a, st, sp, nx, ny, f.spill, hc);
      }
    }
  }
}

// the frame loop's per-particle tail: flood of one finished batch particle (SoilMachine.cpp:292-296)
template <class A> SM_HD void hydro_flood_particle(A& a, const WaterP& p, HydroCount& hc) {
  WFrame st[SM_WSTACK];
  int sp = 0;
  hydro_flood(a, p, 3, st, sp, hc);                                 // spill = 3, water.h:33
  hydro_drain(a, st, sp, hc);
}

// one cell of the full-grid pass WaterParticle::seep(map,...), water.h:335-343
template <class A> SM_HD void hydro_seep_visit(A& a, int x, int y, HydroCount& hc) {
  WFrame st[SM_WSTACK];
  int sp = 0;
  hydro_seep_cell(a, x, y);
  hydro_push(a, st, sp, x, y, 3, hc);
  hydro_drain(a, st, sp, hc);
  hc.cells++;
}

// ------------------------------------------------------------------------------------------------
// Active-cell index for the full-grid pass.
// ------------------------------------------------------------------------------------------------
// Upstream visits every cell in x-major order.  A visit changes nothing unless water is near: seep
// needs a section with saturation != 0 in the column, the water cascade needs an Air-topped cell in the
// 3x3 block.  The device pass therefore visits only cells flagged in a bitmap (one bit per cell in
// visiting order c = x*dimy + y), built by a full-grid classification kernel and kept up to date by the
// executor whenever it makes a cell wet.  Visiting a dry cell is a no-op upstream, and cells are only
// ever added, so the visited sequence is the reference's sequence minus no-ops.
// Levels: level 0 = one bit per cell, level l+1 = one bit per 64-bit word of level l.
#define SM_ACT_LEVELS 5
struct ActiveMap {
  unsigned long long* lvl[SM_ACT_LEVELS];
  unsigned long long nwords[SM_ACT_LEVELS];
  unsigned long long ncells;
  int nlevels;
};
// sizes of the levels for `cells` bits; returns the total number of 64-bit words
static inline unsigned long long active_layout(unsigned long long cells, unsigned long long* nwords, int* nlevels) {
  unsigned long long n = cells, total = 0;
  int l = 0;
  for (;;) {
    n = (n + 63) / 64;
    nwords[l++] = n;
    total += n;
    if (n <= 1 || l == SM_ACT_LEVELS) break;
  }
  *nlevels = l;
  return total;
}
SM_HD void active_set(const ActiveMap& m, unsigned long long idx) {   // single writer
  for (int l = 0; l < m.nlevels; l++) {
    const unsigned long long w = idx >> 6, b = 1ull << (idx & 63);
    const unsigned long long old = m.lvl[l][w];
    if (old & b) return;
    m.lvl[l][w] = old | b;
    if (old) return;                       // the word was already announced upstairs
    idx = w;
  }
}
// smallest flagged index >= from, or ncells
SM_HD unsigned long long active_next(const ActiveMap& m, unsigned long long from) {
  if (from >= m.ncells) return m.ncells;
  int l = 0;
  unsigned long long idx = from;
  for (;;) {
    const unsigned long long w = idx >> 6;
    unsigned long long bits = 0;
    if (w < m.nwords[l]) bits = m.lvl[l][w] & (~0ull << (idx & 63));
    if (bits) {
#if defined(__CUDA_ARCH__)
      const int tz = __ffsll((long long)bits) - 1;
#else
      const int tz = __builtin_ctzll(bits);
#endif
      idx = (w << 6) + (unsigned long long)tz;
      if (l == 0) return idx < m.ncells ? idx : m.ncells;
      l--;
      idx <<= 6;                            // first bit of that word one level down
    } else {
      if (l + 1 >= m.nlevels) return m.ncells;   // nothing left at the top level
      idx = w + 1;                          // continue after this word, one level up
      l++;
    }
  }
}
// flag the 3x3 block around (x, y): an Air-topped cell makes its neighbours' water cascade live
SM_HD void active_mark_block(const ActiveMap& m, int x, int y, int dimx, int dimy) {
  for (int dx = -1; dx <= 1; dx++) {
    const int xx = x + dx;
    if (xx < 0 || xx >= dimx) continue;
    for (int dy = -1; dy <= 1; dy++) {
      const int yy = y + dy;
      if (yy < 0 || yy >= dimy) continue;
      active_set(m, (unsigned long long)xx * dimy + yy);
    }
  }
}
// is this column wet by itself?  (Air on top, or a section that holds water)
template <class A> SM_HD void hydro_classify(A& a, int x, int y, bool& airtop, bool& holds) {
  const Sec32* r = a.rec(x, y);
  airtop = (r->type == SM_AIR);
  holds = false;
  if (r->type == SM_EMPTY) return;
  if (r->saturation != 0.0) { holds = true; return; }
  for (uint32_t b = r->below; b != SM_NIL;) {
    const Sec32 s = a.pool_load(b);
    if (s.saturation != 0.0) { holds = true; return; }
    b = s.below;
  }
}
