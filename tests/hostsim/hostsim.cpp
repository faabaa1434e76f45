// tests/hostsim -- the product's step arithmetic (soilmachine_b200/csrc/sm_core.cuh) compiled
// for the HOST behind the same lockstep driver shape as the oracle.  TEST TOOL ONLY: it lets the
// CPU test-suite check the type-exact transcription against oracle/_ref bit for bit without a
// GPU.  It is not shipped, not linked into the product library and is not a CPU fallback.
#include <array>
#include <memory>
#include <cmath>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include "../../soilmachine_b200/csrc/sm_core.cuh"
#include "../../soilmachine_b200/csrc/sm_noise.cuh"
#include "../../soilmachine_b200/csrc/sm_hydro.cuh"
#include "../../soilmachine_b200/csrc/sm_coop.cuh"
#include "../../soilmachine_b200/csrc/sm_hydro_coop.cuh"
#include "../../soilmachine_b200/csrc/sm_foot.cuh"

namespace {
struct HostMap {
  int dimx = 0, dimy = 0, scale = 80;
  std::vector<Sec32> top, pool;
  std::vector<uint32_t> freelist;
  std::vector<SoilDev> soils;
  std::vector<float> wfreq, wtrack, windfreq;
  int64_t drops = 0;
  double volume_factor = SM_VOLUME_FACTOR;
} M;
ActiveMap* G_act = nullptr;   // seep pass with the active-cell index: cells the executor makes wet are flagged

struct HostAccess {
  int dimx() const { return M.dimx; }
  int dimy() const { return M.dimy; }
  int scale() const { return M.scale; }
  const SoilDev& soil(uint32_t t) const { return M.soils[t]; }
  Sec32* rec(int x, int y) { return &M.top[(size_t)x * M.dimy + y]; }
  double height(int x, int y) { return rec_height(*rec(x, y)); }
  uint32_t surface_of(int x, int y) { return rec_surface(*rec(x, y)); }
  void query(int x, int y, double& h, uint32_t& t) { const Sec32* r = rec(x, y); h = rec_height(*r); t = rec_surface(*r); }
  void begin(int, int) {}
  void target(int, int) {}
  void dirty(int x, int y) { dirty_rec(rec(x, y), x, y); }
  void dirty_rec(Sec32* r, int x, int y) { if (G_act && r->type == SM_AIR) active_mark_block(*G_act, x, y, M.dimx, M.dimy); }
  void wet_mark(int x, int y) { if (G_act) active_set(*G_act, (unsigned long long)x * M.dimy + y); }
  double volume_factor() const { return M.volume_factor; }
  void cascade_prefetch(int, int) {}
  void mark(int) {}
  void note_transfer() {}
  void focus(int, int) {}
  Sec32 pool_load(uint32_t i) { return M.pool[i]; }
  void pool_store(uint32_t i, const Sec32& r) { M.pool[i] = r; }
  uint32_t pool_alloc() {
    if (!M.freelist.empty()) { uint32_t i = M.freelist.back(); M.freelist.pop_back(); return i; }
    M.pool.push_back(Sec32{});
    return (uint32_t)(M.pool.size() - 1);
  }
  void pool_free(uint32_t i) { M.freelist.push_back(i); }
  void track_add(int ind, double v) { M.wtrack[ind] = (float)(M.wtrack[ind] + v); }
  float water_frequency(int ind) { return M.wfreq[ind]; }
  void wind_frequency_touch(int ind) { M.windfreq[ind] = (float)(0.5 * M.windfreq[ind] + 0.5f); }
};

// ---- the warp-cooperative step (sm_coop.cuh) on the host: lanes become loops ---------------------------
// each()/ballot() run the lanes one after the other; the code keeps to "a lane writes only its own slots and
// reads only what earlier phases wrote", so the order of the lanes cannot matter.  G_lane_order reverses it
// to let the tests check exactly that.
int G_lane_order = 0;
struct WarpHost {
  template <class F> void each(int n, F f) {
    if (G_lane_order == 0) for (int l = 0; l < n; l++) f(l);
    else for (int l = n - 1; l >= 0; l--) f(l);
  }
  template <class F> unsigned int ballot(int n, F f) {
    unsigned int m = 0;
    if (G_lane_order == 0) { for (int l = 0; l < n; l++) if (f(l)) m |= 1u << l; }
    else { for (int l = n - 1; l >= 0; l--) if (f(l)) m |= 1u << l; }
    return m;
  }
  template <class F> void one(F f) { f(); }
  bool lead() const { return true; }
};
std::vector<float> G_field_store;
WindField G_field = {nullptr, 0, 0, 0, 0, 0, 0};
// Footprint audit (mode 2): every record access of a cooperative step goes through HostBack::cell_ptr - the staging
// fetches, the in-place accesses beyond the staged blocks, the write-back - so recording the calls by phase shows what
// a step can read or write, and the sets the exact-footprint schedule orders steps by (Foot<KIND>, sm_engine.cu) can
// be checked to contain it.
struct FootAudit {
  bool on = false;
  int phase = 0;                                  // 0 move(), 1 interact(), 2 write-back
  std::vector<std::array<int, 3> > acc;           // x, y, phase
  long long steps = 0, viol[4] = {0, 0, 0, 0};    // accesses outside M / F / the write set / the published box
  long long max_step = 0;                         // largest |npos - ipos|_inf seen
};
FootAudit G_foot;
struct HostBack {   // backing store of CoopWin on the host
  HostAccess h;
  int dimx() const { return M.dimx; }
  int dimy() const { return M.dimy; }
  int scale() const { return M.scale; }
  const SoilDev* soilp(uint32_t t) const { return &M.soils[t]; }
  Sec32* cell_ptr(int x, int y) {
    if (G_foot.on) G_foot.acc.push_back({x, y, G_foot.phase});
    return &M.top[(size_t)x * M.dimy + y];
  }
  void focus(int, int) {}
  Sec32 pool_load(uint32_t i) { return h.pool_load(i); }
  void pool_store(uint32_t i, const Sec32& r) { h.pool_store(i, r); }
  uint32_t pool_alloc() { return h.pool_alloc(); }
  void pool_free(uint32_t i) { h.pool_free(i); }
  float wfreq(int i) const { return M.wfreq[i]; }
  float wtrack(int i) const { return M.wtrack[i]; }
  float windfreq(int i) const { return M.windfreq[i]; }
  void set_wtrack(int i, float v) { M.wtrack[i] = v; }
  void set_windfreq(int i, float v) { M.windfreq[i] = v; }
  void note_transfer() {}
  static constexpr bool kBudget = true;
  static constexpr bool kHydroHooks = true;     // the hooks HostAccess has (active-cell index of the seep pass)
  void air_mark(Sec32* r, int x, int y) { h.dirty_rec(r, x, y); }
  void wet_mark(int x, int y) { h.wet_mark(x, y); }
  double volume_factor() const { return M.volume_factor; }
  void pspeed(float px, float py, double height, float* ps) const { wind_field_pspeed(G_field, px, py, height, ps); }
};
int G_coop = 0;   // 1: the sweeps below run the warp-cooperative step
std::vector<double> BUD;   // coop mode: 6 mass-budget accumulators per particle (sm_coop.cuh)

struct Stats { int64_t steps, sweeps, exit_oob, exit_evap, exit_stall; double seconds; };
// kind 0 water: move() reads plus(ipos); the step touches plus(ipos) U 3x3(npos) and writes {ipos} U 3x3(npos).
// kind 1 wind : move() reads plus(ipos); the step touches and writes 5x5(ipos) U 5x5(npos).   (Foot<KIND>)
void foot_check(int kind, int ix, int iy, bool moved, int nx, int ny, int R) {
  auto ab = [](int v) { return v < 0 ? -v : v; };
  G_foot.steps++;
  if (moved) {
    const long long d = std::max(ab(nx - ix), ab(ny - iy));
    if (d > G_foot.max_step) G_foot.max_step = d;
  }
  for (const auto& a : G_foot.acc) {
    const int dxi = ab(a[0] - ix), dyi = ab(a[1] - iy);
    const int dxn = moved ? ab(a[0] - nx) : 1 << 20, dyn = moved ? ab(a[1] - ny) : 1 << 20;
    const bool plus = dxi + dyi <= 1;
    bool inF, inW;
    if (kind == 0) {
      const bool n3 = dxn <= 1 && dyn <= 1;
      inF = plus || n3;
      inW = (dxi == 0 && dyi == 0) || n3;
    } else {
      inF = (dxi <= 2 && dyi <= 2) || (dxn <= 2 && dyn <= 2);
      inW = inF;
    }
    if (dxi > R || dyi > R) G_foot.viol[3]++;
    if (a[2] == 0 && !plus) G_foot.viol[0]++;
    if (a[2] == 1 && !inF) G_foot.viol[1]++;
    if (a[2] == 2 && !inW) G_foot.viol[2]++;
  }
  G_foot.acc.clear();
}
std::vector<WaterP> W; std::vector<int> Wlive;
std::vector<WindP> D; std::vector<int> Dlive;
}  // namespace

// ---- mode 3: the exact-footprint schedule run ADVERSARIALLY ---------------------------------------------------
// The device lets a step run ahead of lower-index steps whenever the rule of sweep_exact (sm_sweep.cuh) allows it; which
// legal order a GPU run takes depends on timing.  Here one sweep is executed in the legal order that departs most
// from index order: passes over the live particles from the HIGHEST index down, each particle advancing one phase
// (move, then interact) whenever the rule lets it, until all are done - with the step split and staged exactly as on
// the device (a particle's window lives across the two phases while other particles run in between).  The rule is
// stated on cell sets by brute force, not with the closed-form predicates of Foot<KIND>:
//   B < A, B not finished, holds A back
//     before A.move()     while B's (possible) writes can meet plus(ipos_A):
//                           B not moved: its box ipos_B +- R_B;  B moved: {ipos_B} U 3x3(npos_B) (wind: 5x5 U 5x5)
//     before A.interact() while B's (possible) footprint can meet A's footprint plus(ipos_A) U 3x3(npos_A)
//                           (wind: 5x5(ipos_A) U 5x5(npos_A)):  B not moved: its box;  B moved: its footprint.
// If the golden frames come out bit for bit under this order too, footprints that cannot meet commute - the rule is sound.
struct CellSet {
  struct Part { int kind, x, y, r; };        // kind 0: square of half-width r; kind 1: plus
  Part p[2]; int n = 0;
  void rect(int x, int y, int r) { p[n++] = Part{0, x, y, r}; }
  void plus(int x, int y) { p[n++] = Part{1, x, y, 1}; }
  bool has(int x, int y) const {
    for (int i = 0; i < n; i++) {
      const int dx = std::abs(x - p[i].x), dy = std::abs(y - p[i].y);
      if (p[i].kind == 0 ? (dx <= p[i].r && dy <= p[i].r) : (dx + dy <= 1)) return true;
    }
    return false;
  }
  bool meets(const CellSet& o) const {
    for (int i = 0; i < n; i++)
      for (int x = p[i].x - p[i].r; x <= p[i].x + p[i].r; x++)
        for (int y = p[i].y - p[i].r; y <= p[i].y + p[i].r; y++)
          if ((p[i].kind == 0 || std::abs(x - p[i].x) + std::abs(y - p[i].y) <= 1) && o.has(x, y)) return true;
    return false;
  }
};
struct Fly { int id, phase, ix, iy, nx, ny, R; };     // phase 0 not moved, 1 moved, 2 finished
inline int host_reach(const WaterP&) { return 3; }
inline int host_reach(const WindP& p) {                // particle_reach, sm_engine.cu
  const float len = sqrtf(p.sx * p.sx + p.sy * p.sy + p.sz * p.sz);
  int r = (int)floorf(1.0f + 0.8f * len + 0.4f + 0.01f);
  r = r < 1 ? 1 : (r > 3 ? 3 : r);
  return r + 2;
}
template <int KIND> CellSet set_box(const Fly& f) { CellSet s; s.rect(f.ix, f.iy, f.R); return s; }
template <int KIND> CellSet set_M(const Fly& f) { CellSet s; s.plus(f.ix, f.iy); return s; }
template <int KIND> CellSet set_W(const Fly& f) {      // writes of a moved step
  CellSet s;
  if (KIND == 0) { s.rect(f.ix, f.iy, 0); s.rect(f.nx, f.ny, 1); } else { s.rect(f.ix, f.iy, 2); s.rect(f.nx, f.ny, 2); }
  return s;
}
template <int KIND> CellSet set_F(const Fly& f) {      // everything a moved step touches
  CellSet s;
  if (KIND == 0) { s.plus(f.ix, f.iy); s.rect(f.nx, f.ny, 1); } else { s.rect(f.ix, f.iy, 2); s.rect(f.nx, f.ny, 2); }
  return s;
}
long long G_adv_ahead = 0;      // phases executed while a lower-index particle of the sweep was still unfinished
int G_adv_weak = 0;             // negative control: 1 = footprints shrunk by one ring (the test must then FAIL)
inline int adv_move(WarpHost& w, CoopWin<HostBack>& a, WaterP& p, WaterMidCoop& m) { return water_move_coop(w, a, p, m, SM_CW_PLUS); }
inline int adv_move(WarpHost& w, CoopWin<HostBack>& a, WindP& p, WindMidCoop& m) { return wind_move_coop(w, a, p, m, SM_CW_PLUS); }
inline int adv_interact(WarpHost& w, CoopWin<HostBack>& a, WaterP& p, const WaterMidCoop& m) { return water_interact_coop(w, a, p, m); }
inline int adv_interact(WarpHost& w, CoopWin<HostBack>& a, WindP& p, const WindMidCoop& m) { return wind_interact_coop(w, a, p, m); }
// runs one sweep; res[k] = result of live[k]'s step
template <int KIND, class P, class MID>
void sweep_adversarial(std::vector<P>& parts, const std::vector<int>& live, std::vector<int>& res) {
  const size_t n = live.size();
  std::vector<Fly> fly(n);
  std::vector<HostBack> backs(n);
  std::vector<CoopScratch> scr(n);
  std::vector<MID> mids(n);
  std::vector<std::unique_ptr<CoopWin<HostBack> > > wins(n);
  WarpHost w;
  for (size_t k = 0; k < n; k++) {
    const P& p = parts[live[k]];
    fly[k] = Fly{live[k], 0, (int)roundf(p.px), (int)roundf(p.py), 0, 0, host_reach(p)};
    wins[k].reset(new CoopWin<HostBack>(backs[k], &scr[k]));
  }
  res.assign(n, SM_ALIVE);
  size_t left = n;
  for (int pass = 0; left > 0; pass++) {
    if (pass > 100000) { fprintf(stderr, "sweep_adversarial: no progress\n"); abort(); }
    for (size_t kk = n; kk-- > 0;) {          // highest index first (live[] is ascending)
      Fly& A = fly[kk];
      if (A.phase == 2) continue;
      bool ok = true, lower_open = false;
      const CellSet mine = A.phase == 0 ? set_M<KIND>(A) : set_F<KIND>(A);
      for (size_t j = 0; j < kk && ok; j++) {
        const Fly& B = fly[j];
        if (B.phase == 2) continue;
        lower_open = true;
        if (std::abs(B.ix - A.ix) > 16 || std::abs(B.iy - A.iy) > 16) continue;
        CellSet theirs = B.phase == 0 ? set_box<KIND>(B) : (A.phase == 0 ? set_W<KIND>(B) : set_F<KIND>(B));
        if (G_adv_weak)
          for (int q = 0; q < theirs.n; q++) if (theirs.p[q].kind == 0 && theirs.p[q].r > 0) theirs.p[q].r--;
        if (theirs.meets(mine)) ok = false;
      }
      if (!ok) continue;
      if (lower_open) G_adv_ahead++;
      P& p = parts[A.id];
      if (A.phase == 0) {
        const int r = adv_move(w, *wins[kk], p, mids[kk]);
        if (r == SM_ALIVE) { A.nx = (int)roundf(p.px); A.ny = (int)roundf(p.py); A.phase = 1; }
        else { wins[kk]->flush(w); res[kk] = r; A.phase = 2; left--; }
      } else {
        res[kk] = adv_interact(w, *wins[kk], p, mids[kk]);
        wins[kk]->flush(w);
        A.phase = 2; left--;
      }
    }
  }
}

// The closed-form predicates the device schedules use (sm_foot.cuh) against the explicit cell sets, exhaustively over
// every relative position that can occur (|B - A| up to 12, |npos - ipos| up to 2 water / 3 wind, reach 3..5).
// out[2 * k] = cases where the sets meet but predicate k says no (UNSOUND), out[2 * k + 1] = predicate yes but the sets
// do not meet (merely conservative); k = box_hits_M, W_hits_M, F_hits_F, box_hits_F for water, then the same for wind.
template <int KIND> void foot_exhaustive(long long* out) {
  const int S = KIND == 0 ? 2 : 3, half = KIND == 0 ? 1 : 2;
  auto tally = [&](int k, bool pred, bool truth) { if (truth && !pred) out[2 * k]++; if (pred && !truth) out[2 * k + 1]++; };
  for (int RB = 3; RB <= (KIND == 0 ? 3 : 5); RB++)
  for (int bx = -12; bx <= 12; bx++) for (int by = -12; by <= 12; by++) {
    const Fly A0{0, 0, 0, 0, 0, 0, 0}, B0{1, 0, bx, by, 0, 0, RB};
    tally(0, Foot<KIND>::box_hits_M(bx, by, RB), set_box<KIND>(B0).meets(set_M<KIND>(A0)));
    for (int mx = -S; mx <= S; mx++) for (int my = -S; my <= S; my++) {
      Fly B1 = B0; B1.nx = bx + mx; B1.ny = by + my;
      tally(1, Foot<KIND>::W_hits_M(bx, by, B1.nx, B1.ny, 0, 0), set_W<KIND>(B1).meets(set_M<KIND>(A0)));
      Fly A1 = A0; A1.nx = mx; A1.ny = my;      // (reuse the offsets for A's own move)
      tally(3, Foot<KIND>::box_hits_F(0, 0, A1.nx, A1.ny, bx, by, RB), set_box<KIND>(B0).meets(set_F<KIND>(A1)));
      if (RB == 3)
        for (int ax = -S; ax <= S; ax++) for (int ay = -S; ay <= S; ay++) {
          Fly A2 = A0; A2.nx = ax; A2.ny = ay;
          tally(2, Foot<KIND>::F_hits_F(0, 0, ax, ay, bx, by, B1.nx, B1.ny), set_F<KIND>(B1).meets(set_F<KIND>(A2)));
        }
    }
  }
  (void)half;
}

extern "C" {
void hs_init(int dimx, int dimy, int scale, int nsoils, const SoilDev* soils) {
  M = HostMap();
  M.dimx = dimx; M.dimy = dimy; M.scale = scale;
  M.soils.assign(soils, soils + nsoils);
  M.top.resize((size_t)dimx * dimy);
  for (auto& r : M.top) rec_set_empty(r);
  M.wfreq.assign((size_t)dimx * dimy, 0.f); M.wtrack = M.wfreq; M.windfreq = M.wfreq;
}
struct LayerPOD { int32_t type; float min, bias, scale, octaves, lacunarity, gain, frequency; };
// Layermap::initialize (layermap.h:163-216) through the product's noise restatement
void hs_initialize(int seed, int nlayers, const LayerPOD* lay) {
  HostAccess a;
  for (auto& r : M.top) rec_set_empty(r);
  M.pool.clear(); M.freelist.clear();
  for (int l = 0; l < nlayers; l++) {
    LayerDev L{(uint32_t)lay[l].type, lay[l].min, lay[l].bias, lay[l].scale, (int)lay[l].octaves,
               lay[l].lacunarity, lay[l].gain, lay[l].frequency, 0.f};
    L.bounding = fnl_fractal_bounding(L.octaves, L.gain);
    const int zs = layer_zslice(seed, l, nlayers);
    for (int i = 0; i < M.dimx; i++) for (int j = 0; j < M.dimy; j++) {
      double h = layer_value(L, i, j, zs, M.dimx, M.dimy);
      col_add(a, M.top[(size_t)i * M.dimy + j], h, L.type);
    }
  }
}
void hs_set_columns(const int64_t* off, const int32_t* type, const double* size, const double* sat) {
  HostAccess a;
  for (int x = 0; x < M.dimx; x++) for (int y = 0; y < M.dimy; y++) {
    size_t c = (size_t)x * M.dimy + y;
    rec_set_empty(M.top[c]);
    for (int64_t k = off[c]; k < off[c + 1]; k++) col_add(a, M.top[c], size[k], (uint32_t)type[k], sat ? sat[k] : 0.0);
  }
}
int64_t hs_nsections() {
  int64_t n = 0;
  for (auto& r : M.top) { if (r.type == SM_EMPTY) continue; n++; for (uint32_t b = r.below; b != SM_NIL; b = M.pool[b].below) n++; }
  return n;
}
void hs_get_columns(int64_t* off, int32_t* type, double* size, double* floor_, double* sat) {
  int64_t n = 0;
  for (size_t c = 0; c < M.top.size(); c++) {
    off[c] = n;
    std::vector<Sec32> st;
    const Sec32& r = M.top[c];
    if (r.type != SM_EMPTY) { st.push_back(r); for (uint32_t b = r.below; b != SM_NIL; b = M.pool[b].below) st.push_back(M.pool[b]); }
    for (size_t i = st.size(); i-- > 0;) { type[n] = (int32_t)st[i].type; size[n] = st[i].size; floor_[n] = st[i].floor; sat[n] = st[i].saturation; n++; }
  }
  off[M.top.size()] = n;
}
void hs_heights(double* out) { for (size_t c = 0; c < M.top.size(); c++) out[c] = rec_height(M.top[c]); }
void hs_get_frequency(float* a, float* b, float* c) {
  size_t n = M.wfreq.size() * 4;
  if (a) memcpy(a, M.wfreq.data(), n); if (b) memcpy(b, M.wtrack.data(), n); if (c) memcpy(c, M.windfreq.data(), n);
}
void hs_set_frequency(const float* a, const float* b, const float* c) {
  size_t n = M.wfreq.size() * 4;
  if (a) memcpy(M.wfreq.data(), a, n); if (b) memcpy(M.wtrack.data(), b, n); if (c) memcpy(M.windfreq.data(), c, n);
}
void hs_normal(int x, int y, float* o) { HostAccess a; sm_f3 n = map_normal(a, x, y); o[0] = n.x; o[1] = n.y; o[2] = n.z; }
double hs_height_f(float x, float y) { HostAccess a; return map_height_bilinear(a, x, y); }
void hs_add(int x, int y, double s, int t) { HostAccess a; col_add(a, *a.rec(x, y), s, (uint32_t)t); }
double hs_remove(int x, int y, double h) { HostAccess a; return col_remove(a, *a.rec(x, y), h); }
void hs_cascade(float x, float y, int loop) { HostAccess a; Cascade<3, HostAccess>::run(a, (int)roundf(x), (int)roundf(y), loop); }

void hs_set_mode(int coop, int lane_order) { G_coop = coop; G_lane_order = lane_order; }
// footprint audit of the steps run in mode 2 since the last call: steps, accesses outside plus(ipos) during move(),
// outside the footprint during interact(), write-backs outside the write set, largest |npos - ipos|
void hs_check_foot_predicates(long long* out16) {
  for (int i = 0; i < 16; i++) out16[i] = 0;
  foot_exhaustive<0>(out16);
  foot_exhaustive<1>(out16 + 8);
}
void hs_adversarial_weaken(int on) { G_adv_weak = on; }
long long hs_adversarial_ahead(void) { const long long v = G_adv_ahead; G_adv_ahead = 0; return v; }
void hs_footprint_audit(long long* out6) {
  out6[0] = G_foot.steps; out6[1] = G_foot.viol[0]; out6[2] = G_foot.viol[1]; out6[3] = G_foot.viol[2]; out6[4] = G_foot.max_step;
  out6[5] = G_foot.viol[3];
  G_foot = FootAudit();
}
void hs_set_volume_factor(double v) { M.volume_factor = v; }
// attach (v4 != null) or detach a lattice velocity field for the wind particles' prevailing wind
void hs_set_wind_field(const float* v4, int nx, int ny, int nz) {
  if (!v4) { G_field = WindField{nullptr, 0, 0, 0, 0, 0, 0}; return; }
  G_field_store.assign(v4, v4 + (size_t)nx * ny * nz * 4);
  G_field = WindField{G_field_store.data(), nx, ny, nz, M.dimx, M.dimy, M.scale};
}
void hs_budget(double* per_particle) { memcpy(per_particle, BUD.data(), BUD.size() * sizeof(double)); }
void hs_water_begin(int n, const float* xy) {
  HostAccess a; W.clear(); Wlive.clear(); BUD.assign((size_t)n * SM_BUDGET_SLOTS, 0.0);
  for (int i = 0; i < n; i++) {
    WaterP p{xy[2 * i], xy[2 * i + 1], 0.f, 0.f, 1.0, 0.0, 0};
    p.contains = spawn_contains(a, p.px, p.py);
    W.push_back(p); Wlive.push_back(i);
  }
}
int hs_water_sweep(Stats* st) {
  HostAccess a; std::vector<int> next;
  std::vector<int> adv;
  if (G_coop == 3) sweep_adversarial<0, WaterP, WaterMidCoop>(W, Wlive, adv);
  size_t kpos = 0;
  for (int i : Wlive) {
    int r;
    if (G_coop == 3) r = adv[kpos++];
    else
    if (G_coop) {
      WarpHost w; HostBack b; CoopScratch sc; CoopWin<HostBack> cw(b, &sc);
      if (G_coop == 2) {        // staged and split as the exact-footprint schedule does (sweep_exact), with the audit
        const int ix = (int)roundf(W[i].px), iy = (int)roundf(W[i].py), R = host_reach(W[i]);
        WaterMidCoop mid;
        G_foot.on = true; G_foot.phase = 0;
        r = water_move_coop(w, cw, W[i], mid, SM_CW_PLUS);
        const bool moved = r == SM_ALIVE;
        const int nx = (int)roundf(W[i].px), ny = (int)roundf(W[i].py);
        G_foot.phase = 1;
        if (moved) r = water_interact_coop(w, cw, W[i], mid);
        G_foot.phase = 2;
        cw.flush(w);
        G_foot.on = false;
        foot_check(0, ix, iy, moved, nx, ny, R);
      } else {
        r = water_step_coop(w, cw, W[i]);
        cw.flush(w);
      }
      for (int k = 0; k < SM_BUDGET_SLOTS; k++) BUD[(size_t)i * SM_BUDGET_SLOTS + k] += sc.acc[k];
    } else r = water_step(a, W[i]);
    if (r == SM_EXIT_OOB) { st->exit_oob++; continue; }
    if (r == SM_EXIT_STALL) { st->exit_stall++; continue; }
    st->steps++;
    if (r == SM_EXIT_EVAP) { st->exit_evap++; continue; }
    next.push_back(i);
  }
  Wlive.swap(next); st->sweeps++;
  return (int)Wlive.size();
}
void hs_water_state(float* pos, float* speed, double* vol, double* sed, int32_t* cont, int32_t* alive) {
  for (size_t i = 0; i < W.size(); i++) {
    pos[2 * i] = W[i].px; pos[2 * i + 1] = W[i].py; speed[2 * i] = W[i].sx; speed[2 * i + 1] = W[i].sy;
    vol[i] = W[i].volume; sed[i] = W[i].sediment; cont[i] = (int32_t)W[i].contains; alive[i] = 0;
  }
  for (int i : Wlive) alive[i] = 1;
}
void hs_wind_begin(int n, const float* xy) {
  HostAccess a; D.clear(); Dlive.clear(); BUD.assign((size_t)n * SM_BUDGET_SLOTS, 0.0);
  for (int i = 0; i < n; i++) {
    WindP p{xy[2 * i], xy[2 * i + 1], -2.f, 0.f, 1.f, 0.0, 0.0, 0};
    p.contains = spawn_contains(a, p.px, p.py);
    D.push_back(p); Dlive.push_back(i);
  }
}
int hs_wind_sweep(Stats* st) {
  HostAccess a; std::vector<int> next;
  std::vector<int> adv;
  if (G_coop == 3) sweep_adversarial<1, WindP, WindMidCoop>(D, Dlive, adv);
  size_t kpos = 0;
  for (int i : Dlive) {
    int r;
    if (G_coop == 3) r = adv[kpos++];
    else
    if (G_coop) {
      WarpHost w; HostBack b; CoopScratch sc; CoopWin<HostBack> cw(b, &sc);
      if (G_coop == 2) {
        const int ix = (int)roundf(D[i].px), iy = (int)roundf(D[i].py), R = host_reach(D[i]);
        WindMidCoop mid;
        G_foot.on = true; G_foot.phase = 0;
        r = wind_move_coop(w, cw, D[i], mid, SM_CW_PLUS);
        const bool moved = r == SM_ALIVE;
        const int nx = (int)roundf(D[i].px), ny = (int)roundf(D[i].py);
        G_foot.phase = 1;
        if (moved) r = wind_interact_coop(w, cw, D[i], mid);
        G_foot.phase = 2;
        cw.flush(w);
        G_foot.on = false;
        foot_check(1, ix, iy, moved, nx, ny, R);
      } else {
        r = wind_step_coop(w, cw, D[i]);
        cw.flush(w);
      }
      for (int k = 0; k < SM_BUDGET_SLOTS; k++) BUD[(size_t)i * SM_BUDGET_SLOTS + k] += sc.acc[k];
    } else r = wind_step(a, D[i]);
    if (r != SM_ALIVE) { st->exit_oob++; continue; }
    st->steps++;
    next.push_back(i);
  }
  Dlive.swap(next); st->sweeps++;
  return (int)Dlive.size();
}
void hs_wind_state(float* pos, float* speed3, double* h, double* sed, int32_t* cont, int32_t* alive) {
  for (size_t i = 0; i < D.size(); i++) {
    pos[2 * i] = D[i].px; pos[2 * i + 1] = D[i].py;
    speed3[3 * i] = D[i].sx; speed3[3 * i + 1] = D[i].sy; speed3[3 * i + 2] = D[i].sz;
    h[i] = D[i].height; sed[i] = D[i].sediment; cont[i] = (int32_t)D[i].contains; alive[i] = 0;
  }
  for (int i : Dlive) alive[i] = 1;
}

// ---- pooling hydrology (sm_hydro.cuh) ------------------------------------------------------------
void hs_water_flood(HydroCount* out) {
  HostAccess a; HydroCount hc{};
  std::vector<char> live(W.size(), 0);
  for (int i : Wlive) live[i] = 1;
  if (G_coop) {            // the warp-cooperative executor (sm_hydro_coop.cuh), lanes as loops
    WarpHost w; HostBack b; CoopScratch sc; HydroScratch hx; CoopWin<HostBack> cw(b, &sc);
    for (size_t i = 0; i < W.size(); i++) if (!live[i]) hydro_flood_particle_coop(w, cw, &hx, W[i], hc);
  } else {
    for (size_t i = 0; i < W.size(); i++) if (!live[i]) hydro_flood_particle(a, W[i], hc);
  }
  if (out) *out = hc;
}
// mode 0: visit every cell in x-major order, as upstream; mode 1: classify + visit the flagged cells only,
// the way the device pass does
void hs_seep(int mode, HydroCount* out) {
  HostAccess a; HydroCount hc{};
  WarpHost w; HostBack b; CoopScratch sc; HydroScratch hx; CoopWin<HostBack> cw(b, &sc);
  auto visit = [&](int x, int y) {
    if (G_coop) hydro_seep_visit_coop(w, cw, &hx, x, y, hc);
    else hydro_seep_visit(a, x, y, hc);
  };
  if (mode == 0) {
    for (int x = 0; x < M.dimx; x++) for (int y = 0; y < M.dimy; y++) visit(x, y);
  } else {
    ActiveMap am{};
    const unsigned long long cells = (unsigned long long)M.dimx * M.dimy;
    unsigned long long total = active_layout(cells, am.nwords, &am.nlevels);
    std::vector<unsigned long long> store(total, 0ull);
    unsigned long long off = 0;
    for (int l = 0; l < am.nlevels; l++) { am.lvl[l] = store.data() + off; off += am.nwords[l]; }
    am.ncells = cells;
    for (int x = 0; x < M.dimx; x++) for (int y = 0; y < M.dimy; y++) {
      bool airtop, holds;
      hydro_classify(a, x, y, airtop, holds);
      if (airtop) active_mark_block(am, x, y, M.dimx, M.dimy);
      if (holds) active_set(am, (unsigned long long)x * M.dimy + y);
    }
    G_act = &am;
    for (unsigned long long c = active_next(am, 0); c < cells; c = active_next(am, c + 1))
      visit((int)(c / M.dimy), (int)(c % M.dimy));
    G_act = nullptr;
  }
  if (out) *out = hc;
}
// enumerate a bitmap built from `idx` (n distinct cell numbers) with active_next; returns the count written
int hs_active_selftest(unsigned long long cells, const unsigned long long* idx, int n, unsigned long long* out) {
  ActiveMap am{};
  unsigned long long total = active_layout(cells, am.nwords, &am.nlevels);
  std::vector<unsigned long long> store(total, 0ull);
  unsigned long long off = 0;
  for (int l = 0; l < am.nlevels; l++) { am.lvl[l] = store.data() + off; off += am.nwords[l]; }
  am.ncells = cells;
  for (int i = n - 1; i >= 0; i--) { active_set(am, idx[i]); active_set(am, idx[i]); }
  int k = 0;
  for (unsigned long long c = active_next(am, 0); c < cells; c = active_next(am, c + 1)) { if (k <= n) out[k] = c; k++; }
  return k;
}
}
