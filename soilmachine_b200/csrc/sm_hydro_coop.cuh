// sm_hydro_coop.cuh -- pooling hydrology (flood, water-table cascade, seep) executed by ONE WARP.
//
// Same canonical order and the same arithmetic as sm_hydro.cuh (reference: water.h:123-343): floods in ascending
// particle index, the seep pass over the flagged cells in x-major order, every flood / visit atomic with its
// nested particles.  That order is sequential by definition, so what a warp can shorten is the work inside one
// flood or visit:
//   * a water-cascade frame evaluates its (up to) eight neighbours on eight lanes at once; the first neighbour in
//     visiting order that acts is executed, the ones before it had no effect (the early exits of water.h:211-234
//     are side-effect free), the ones after it are evaluated again on the new state - exactly the states the
//     reference loop sees (same argument as CascadeCoop in sm_coop.cuh);
//   * opening a frame (eight heights + stable rank sort) is lane-parallel;
//   * a nested particle runs the cooperative particle-step (water_step_coop), the flood's terrain cascade the
//     cooperative cascade.
// Column mutations are done by one lane.  Written against the same two policies as sm_coop.cuh (warp W, accessor
// A = CoopWin<B>), so tests/hostsim runs this very code on the host against the reference's golden vectors.
#pragma once
#include "sm_coop.cuh"
#include "sm_hydro.cuh"

struct HydroScratch {          // per warp; written by one lane, read by all
  WFrame st[SM_WSTACK];
};

// gather the in-bounds neighbours and their visiting order (water.h:155-183): highest first, stable
template <class W, class A> SM_HD void hydro_open_coop(W& w, A& a, WFrame* f, int cx, int cy, int spill) {
  const int dimx = a.dimx(), dimy = a.dimy();
  double* const hs = a.s->hs[0];
  unsigned char* const ord = a.s->ord[0];
  const unsigned int inb = ((cx > 0 ? 0x07u : 0u) | 0x18u | (cx < dimx - 1 ? 0xE0u : 0u)) &
                           ((cy > 0 ? 0x29u : 0u) | 0x42u | (cy < dimy - 1 ? 0x94u : 0u));   // :171-172
  w.each(8, [&](int k) {
    const int kk = k + (k >= 4 ? 1 : 0);
    hs[k] = ((inb >> k) & 1u) ? a.height(cx + kk / 3 - 1, cy + kk % 3 - 1) : -1.0e300;
  });
  w.each(8, [&](int k) {
    const double hk = hs[k];
    int r = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const double hj = hs[j];
      r += (j > k) ? (hj > hk ? 1 : 0) : ((j < k) ? (hk > hj ? 0 : 1) : 0);
    }
    ord[r] = (unsigned char)k;
  });
  w.one([&]() {
    unsigned int order = 0;
    for (int r = 0; r < 8; r++) order |= (unsigned int)ord[r] << (4 * r);
    f->cx = cx; f->cy = cy; f->order = order; f->num = SM_POPC(inb); f->i = 0; f->spill = spill;
  });
}
template <class W, class A>
SM_HD void hydro_push_coop(W& w, A& a, HydroScratch* hsx, int& sp, int cx, int cy, int spill, HydroCount& hc) {
  if (sp >= SM_WSTACK) { hc.overflow++; return; }
  hydro_open_coop(w, a, &hsx->st[sp], cx, cy, spill);
  sp++;
}

// would the water cascade act on neighbour (nx, ny) of (cx, cy) right now?  (water.h:185-234; no side effects)
template <class A> SM_HD bool hydro_neighbour_acts(A& a, int cx, int cy, int nx, int ny, int SCALE) {
  const Sec32* pa = a.rec(cx, cy);
  const Sec32* pb = a.rec(nx, ny);
  double whA = 0, whB = 0, fA = 0.0, fB = 0.0;
  if (pa->type != SM_EMPTY) { whA = pa->size; fA = pa->floor; }
  if (pb->type != SM_EMPTY) { whB = pb->size; fB = pb->floor; }
  const double num = (fA + whA - fB - whB) * (double)SCALE;
  if (num == 0) return false;
  const Sec32* top = (num > 0) ? pa : pb;
  if (top->type != SM_AIR) return false;
  const double diff = num / 80.0;
  if (diff == 0) return false;
  double transfer = fabs(diff) / 2.0;
  const double wh = top->size;
  transfer = (wh < transfer) ? wh : transfer;
  return !(transfer <= 0);
}

template <class W, class A>
SM_HD void hydro_flood_coop(W& w, A& a, const WaterP& p, int spill, HydroScratch* hsx, int& sp, HydroCount& hc);

// run every open frame to its end (water.h:185-281 as a frame machine, see sm_hydro.cuh)
template <class W, class A> SM_HD void hydro_drain_coop(W& w, A& a, HydroScratch* hsx, int& sp, HydroCount& hc) {
  const int SCALE = a.scale();
  SM_UNROLL1
  while (sp > 0) {
    WFrame* const fp = &hsx->st[sp - 1];
    const int cx = fp->cx, cy = fp->cy, fnum = fp->num, fi = fp->i;
    const unsigned int order = fp->order;
    if (fi >= fnum) { sp--; continue; }
    const unsigned int act = w.ballot(fnum, [&](int r) {
      if (r < fi) return false;
      const int k = (int)((order >> (4 * r)) & 7u);
      const int kk = k + (k >= 4 ? 1 : 0);
      return hydro_neighbour_acts(a, cx, cy, cx + kk / 3 - 1, cy + kk % 3 - 1, SCALE);
    });
    if (act == 0) { sp--; continue; }            // the rest of the frame is a no-op
    const int i = SM_FFS(act) - 1;
    const int k = (int)((order >> (4 * i)) & 7u);
    const int kk = k + (k >= 4 ? 1 : 0);
    const int nx = cx + kk / 3 - 1, ny = cy + kk % 3 - 1;
    Sec32* const pa = a.rec(cx, cy);
    Sec32* const pb = a.rec(nx, ny);
    // the acting neighbour, evaluated again by every lane (uniform): same expressions as hydro_drain
    double whA = 0, whB = 0, fA = 0.0, fB = 0.0;
    if (pa->type != SM_EMPTY) { whA = pa->size; fA = pa->floor; }
    if (pb->type != SM_EMPTY) { whB = pb->size; fB = pb->floor; }
    const double num = (fA + whA - fB - whB) * (double)SCALE;
    Sec32* const top = (num > 0) ? pa : pb;
    Sec32* const bot = (num > 0) ? pb : pa;
    const int tx = (num > 0) ? cx : nx, ty = (num > 0) ? cy : ny;
    const int bx = (num > 0) ? nx : cx, by = (num > 0) ? ny : cy;
    const double diff = num / 80.0;
    double transfer = fabs(diff) / 2.0;                             // :227
    const double wh = top->size;                                    // :230
    transfer = (wh < transfer) ? wh : transfer;
    const int fspill = fp->spill;
    if (transfer == wh) {                                           // :240-258 all of it leaves as a particle
      w.one([&]() {
        fp->i = i + 1;
        a.focus(tx, ty);
        col_remove(a, *top, transfer);
        a.dirty_rec(top, tx, ty);
      });
      WaterP q;
      q.px = (float)tx; q.py = (float)ty;
      {
        const float dx = (float)bx - (float)tx, dy = (float)by - (float)ty;   // :246
        const float inv = 1.0f / sqrtf(dx * dx + dy * dy);
        q.sx = SM_SQRT2F * (dx * inv);
        q.sy = SM_SQRT2F * (dy * inv);
      }
      q.volume = transfer / a.volume_factor();                      // :250
      q.sediment = 0.0;
      q.contains = a.soil(rec_surface(*top)).transports;            // see sm_hydro.cuh: the ctor's rand() cannot reach the map
      hc.nested++;
      SM_UNROLL1
      for (;;) {                                                    // :252-253
        const int rc = water_step_coop(w, a, q);
        // Air-topped records the step modified keep the seep pass's index up to date (one lane), then write back
        const uint32_t dm = a.dirtym;
        w.one([&]() {
          for (int l = 0; l < SM_CW_SLOTS; l++)
            if ((dm >> l) & 1u) {
              const int ox = l < 9 ? a.ax : a.bx, oy = l < 9 ? a.ay : a.by, kq = l < 9 ? l : l - 9;
              a.b.air_mark(&a.s->win[l], ox + kq / 3 - 1, oy + kq % 3 - 1);
            }
        });
        a.flush(w);
        if (rc == SM_ALIVE || rc == SM_EXIT_EVAP) hc.nested_steps++;
        if (rc != SM_ALIVE) break;
      }
      a.detach();
      hydro_flood_coop(w, a, q, fspill, hsx, sp, hc);               // :254
    } else {                                                        // :260-272
      w.one([&]() {
        fp->i = i + 1;
        a.focus(tx, ty);
        col_remove(a, *top, transfer);
        a.dirty_rec(top, tx, ty);
        a.focus(bx, by);
        col_add(a, *bot, transfer, SM_AIR);
        if (bot->type != SM_EMPTY) bot->saturation = 1.0;           // map.top(bpos)->saturation = 1.0f
        a.wet_mark(bx, by);
        a.dirty_rec(bot, bx, by);
        if (fspill > 0) fp->spill = fspill - 1;                     // :277-278 cascade(npos, --spill)
      });
      hc.transfers++;
      if (fspill > 0) hydro_push_coop(w, a, hsx, sp, nx, ny, fspill - 1, hc);
    }
  }
}

// WaterParticle::flood, water.h:123-145; its trailing cascade call becomes a pushed frame
template <class W, class A>
SM_HD void hydro_flood_coop(W& w, A& a, const WaterP& p, int spill, HydroScratch* hsx, int& sp, HydroCount& hc) {
  if (p.volume < SM_MINVOL || spill-- <= 0) return;                 // :125-126
  hc.floods++;
  const int ix = (int)p.px, iy = (int)p.py;                         // :128 ipos = pos truncates
  Sec32* const r = a.rec(ix, iy);
  const double sed = p.sediment * a.soil(p.contains).equrate;
  const uint32_t what = p.contains;
  w.one([&]() {
    a.focus(ix, iy);
    col_add(a, *r, sed, what);                                      // :133
    a.dirty_rec(r, ix, iy);
  });
  CascadeCoop<0, W, A>::run(w, a, (int)roundf(p.px), (int)roundf(p.py), 0);   // :134
  const double water = p.volume * a.volume_factor();
  w.one([&]() {
    a.focus(ix, iy);
    col_add(a, *r, water, SM_AIR);                                  // :138
    a.dirty_rec(r, ix, iy);
    hydro_seep_cell(a, ix, iy);                                     // :139
  });
  hydro_push_coop(w, a, hsx, sp, ix, iy, spill, hc);                // :140
}

// the frame loop's per-particle tail: flood of one finished batch particle (SoilMachine.cpp:292-296)
template <class W, class A>
SM_HD void hydro_flood_particle_coop(W& w, A& a, HydroScratch* hsx, const WaterP& p, HydroCount& hc) {
  int sp = 0;
  a.detach();
  hydro_flood_coop(w, a, p, 3, hsx, sp, hc);                        // spill = 3, water.h:33
  hydro_drain_coop(w, a, hsx, sp, hc);
}
// one cell of the full-grid pass WaterParticle::seep(map,...), water.h:335-343
template <class W, class A>
SM_HD void hydro_seep_visit_coop(W& w, A& a, HydroScratch* hsx, int x, int y, HydroCount& hc) {
  int sp = 0;
  a.detach();
  w.one([&]() { hydro_seep_cell(a, x, y); });
  hydro_push_coop(w, a, hsx, sp, x, y, 3, hc);
  hydro_drain_coop(w, a, hsx, sp, hc);
  hc.cells++;
}
