// sm_lbm.cuh -- the wind field: D3Q19 lattice Boltzmann, TRT collision.
//
// Reference: source/include/lbmwind/lbmwind.h:75-236 (buffers, dispatch order), shader/LBM/lbm.cs (velocity set,
// equilibrium), init.cs:7-24, collide.cs:10-58, stream.cs:7-38, shader/move.cs:23-52 (tracer advection);
// boundary from the terrain SoilMachine.cpp:234-239.  Upstream keeps this a visual toy - WindParticle uses a
// constant prevailing wind (wind.h:29) - and runs it as OpenGL compute shaders, three dispatches per step over
// an array-of-structures buffer F[cell*19 + q].
//
// Here: structure of arrays F[q][cell] (every load and store of a warp is one contiguous 128-byte line), fp32,
// and ONE kernel per step - collide in registers, then push the 19 post-collision populations straight to the
// neighbours' slots of the other buffer (ping-pong).  Algorithmic traffic per cell and step: 19 x 4 B read +
// 19 x 4 B written (+ 4 B boundary flag in, 20 B density/velocity out) = 176 B: bandwidth-bound, no reuse.
//   * the five driven faces (y = NY-1, x = 0, x = NX-1, z = 0, z = NZ-1) are rewritten to the forcing equilibrium
//     after the push (stream.cs:25-35): the owner thread writes them, pushers skip such destinations;
//   * populations that would arrive from outside the domain are never written upstream and keep their previous
//     value: the owner thread carries them over to the other buffer.
// Arithmetic mirrors oracle/lbm_oracle.c statement by statement (fp32, left to right, no contraction); parity with
// upstream's GLSL is unpinned (no GL here, no reference vectors) and that file is the definition.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <utility>

#define LBM_Q 19
// velocity set (lbm.cs:56-80).  Compile-time tables: after unrolling every c is a literal, so terms with c = 0
// vanish and c = +-1 become an add / subtract.  That is exact: x*1 = x, x*(-1) = -x, and a +-0 term can only flip
// the sign of an exact zero, which no later operation of the scheme can see (every population is positive and
// every sum it enters is non-zero).
struct LbmSet {
  static constexpr int c[LBM_Q][3] = {
      {0, 0, 0},
      {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1},
      {1, 1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, -1}, {0, 1, 1}, {0, -1, -1},
      {1, -1, 0}, {-1, 1, 0}, {1, 0, -1}, {-1, 0, 1}, {0, 1, -1}, {0, -1, 1}};
  static constexpr int cp[LBM_Q] = {0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15, 18, 17};
  static constexpr int wclass[LBM_Q] = {0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2};   // 1/3, 1/18, 1/36
};
struct LbmConst {
  float w[3];              // 1/3, 1/18, 1/36
  float force[3];          // 0.05 * (-2, 0, 1), lbm.cs:35
  float cs2, cs4;          // 1/cs/cs, 1/cs/cs/cs/cs with cs = 1/sqrt(3), lbm.cs:82-84 (evaluated in fp32, left to right)
  float eq_force[LBM_Q];   // equilibrium(q, 1, force)
  float eq_rest[LBM_Q];    // equilibrium(q, 1, 0)
};
__constant__ LbmConst c_lbm;

struct LbmDev {
  int nx, ny, nz;
  float* F[2];       // [q][cell]
  float* B;          // boundary flag per cell (> 0: obstacle)
  float* RHO;
  float4* V;
};

// compile-time loop over the populations: f(integral_constant<int, q>) for q = 0..18
template <class F, int... Is>
__host__ __device__ __forceinline__ void lbm_for_q_impl(F&& f, std::integer_sequence<int, Is...>) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <class F> __host__ __device__ __forceinline__ void lbm_for_q(F&& f) {
  lbm_for_q_impl(f, std::make_integer_sequence<int, LBM_Q>{});
}
#define LBM_QC(qc) decltype(qc)::value

// dot(v, c[q]) = v.x*c.x + v.y*c.y + v.z*c.z, left to right, zero terms dropped
template <int Q_> __host__ __device__ __forceinline__ float lbm_dot(const float* v) {
  constexpr int cx = LbmSet::c[Q_][0], cy = LbmSet::c[Q_][1], cz = LbmSet::c[Q_][2];
  float d = 0.0f;
  bool any = false;
  if (cx != 0) { d = (cx > 0) ? v[0] : -v[0]; any = true; }
  if (cy != 0) { const float t = (cy > 0) ? v[1] : -v[1]; d = any ? d + t : t; any = true; }
  if (cz != 0) { const float t = (cz > 0) ? v[2] : -v[2]; d = any ? d + t : t; any = true; }
  return d;
}
// equilibrium(q, rho, v), lbm.cs:88-97
template <int Q_> __host__ __device__ __forceinline__ float lbm_equilibrium(const LbmConst& K, float rho, const float* v) {
  const float wq = K.w[LbmSet::wclass[Q_]];
  const float d = lbm_dot<Q_>(v);
  float eq = 0.0f;
  eq += wq * rho;
  eq += wq * rho * d * K.cs2;
  eq += wq * rho * (d * d) * 0.5f * K.cs4;
  eq -= wq * rho * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) * 0.5f * K.cs2;
  return eq;
}
template <int Q_> struct LbmEqAll {     // feq[0..Q_] for one cell
  static __host__ __device__ __forceinline__ void run(const LbmConst& K, float rho, const float* v, float* feq) {
    LbmEqAll<Q_ - 1>::run(K, rho, v, feq);
    feq[Q_] = lbm_equilibrium<Q_>(K, rho, v);
  }
};
template <> struct LbmEqAll<-1> {
  static __host__ __device__ __forceinline__ void run(const LbmConst&, float, const float*, float*) {}
};

// init.cs:7-24
__global__ void __launch_bounds__(256) k_lbm_init(LbmDev L, int buf) {
  const size_t n = (size_t)L.nx * L.ny * L.nz;
  for (size_t ind = (size_t)blockIdx.x * blockDim.x + threadIdx.x; ind < n; ind += (size_t)gridDim.x * blockDim.x) {
    const bool solid = L.B[ind] > 0;
    float rho = 0.0f, v[3] = {0.0f, 0.0f, 0.0f};
    lbm_for_q([&](auto qc) {
      constexpr int q = LBM_QC(qc);
      const float f = solid ? c_lbm.eq_rest[q] : c_lbm.eq_force[q];
      L.F[buf][(size_t)q * n + ind] = f;
      rho += f;
      constexpr int cx = LbmSet::c[q][0], cy = LbmSet::c[q][1], cz = LbmSet::c[q][2];
      if (cx > 0) v[0] += f; else if (cx < 0) v[0] -= f;
      if (cy > 0) v[1] += f; else if (cy < 0) v[1] -= f;
      if (cz > 0) v[2] += f; else if (cz < 0) v[2] -= f;
    });
    L.RHO[ind] = rho;
    L.V[ind] = make_float4(v[0] / rho, v[1] / rho, v[2] / rho, 1.0f);
  }
}

// collide.cs:10-58 + stream.cs:7-38 in one pass: read buffer `src`, write buffer `src ^ 1`
__global__ void __launch_bounds__(256) k_lbm_step(LbmDev L, int src) {
  const size_t n = (size_t)L.nx * L.ny * L.nz;
  const float* __restrict__ Fa = L.F[src];
  float* __restrict__ Fb = L.F[src ^ 1];
  const float tau = 0.56f, dt = 1.0f;
  const float omega_plus = 1.0f / tau;
  const float lambda = 0.25f;
  const float omega_minus = 1.0f / (lambda / (1.0f / omega_plus - 0.5f) + 0.5f);
  for (size_t ind = (size_t)blockIdx.x * blockDim.x + threadIdx.x; ind < n; ind += (size_t)gridDim.x * blockDim.x) {
    const int z = (int)(ind % L.nz), y = (int)((ind / L.nz) % L.ny), x = (int)(ind / ((size_t)L.nz * L.ny));
    float f[LBM_Q];
#pragma unroll
    for (int q = 0; q < LBM_Q; q++) f[q] = Fa[(unsigned int)q * (unsigned int)n + (unsigned int)ind];   // 19 n < 2^31 (sm_lbm_create)
    const bool solid = L.B[ind] > 0.0f;
    // ---- collide ----
    float rho = 0.0f;                                           // getRho, lbm.cs:101-108
#pragma unroll
    for (int q = 0; q < LBM_Q; q++) rho += f[q];
    float v[3] = {0.0f, 0.0f, 0.0f};                            // getV, lbm.cs:112-119
    lbm_for_q([&](auto qc) {
      constexpr int q = LBM_QC(qc);
      constexpr int cx = LbmSet::c[q][0], cy = LbmSet::c[q][1], cz = LbmSet::c[q][2];
      if (cx > 0) v[0] += f[q]; else if (cx < 0) v[0] -= f[q];
      if (cy > 0) v[1] += f[q]; else if (cy < 0) v[1] -= f[q];
      if (cz > 0) v[2] += f[q]; else if (cz < 0) v[2] -= f[q];
    });
#pragma unroll
    for (int k = 0; k < 3; k++) v[k] = v[k] / rho;
    v[1] += dt * 0.0001f * -1.0f / (2.0f * rho);                // gravity along c[4] = (0,-1,0), collide.cs:21
    L.RHO[ind] = rho;
    L.V[ind] = make_float4(v[0], v[1], v[2], 0.0f);
    float feq[LBM_Q];
    LbmEqAll<LBM_Q - 1>::run(c_lbm, rho, v, feq);
    float post[LBM_Q];
    lbm_for_q([&](auto qc) {                                    // TRT, collide.cs:33-56
      constexpr int q = LBM_QC(qc);
      constexpr int o = LbmSet::cp[q];
      const float f_plus = 0.5f * (f[q] + f[o]);
      const float f_minus = 0.5f * (f[q] - f[o]);
      const float feq_plus = 0.5f * (feq[q] + feq[o]);
      const float feq_minus = 0.5f * (feq[q] - feq[o]);
      post[q] = f[q] - omega_plus * (f_plus - feq_plus) - omega_minus * (f_minus - feq_minus);
      if (solid) post[q] = c_lbm.eq_rest[q];
    });
    // ---- stream (push) ----
    // Where a population goes / comes from is decided per axis by compile-time offsets and nine per-cell flags,
    // so each of the 19 slots costs a couple of predicate operations instead of a dozen comparisons.
    const bool x0 = (x == 0), x1 = (x == 1), xl = (x == L.nx - 1), xl1 = (x == L.nx - 2);
    const bool y0 = (y == 0), yl = (y == L.ny - 1), yl1 = (y == L.ny - 2);
    const bool z0 = (z == 0), z1 = (z == 1), zl = (z == L.nz - 1), zl1 = (z == L.nz - 2);
    const bool driven = (yl || x0 || xl || z0 || zl);
    const unsigned int cell = (unsigned int)ind;
    const unsigned int un = (unsigned int)n;
    const int sxy = L.ny * L.nz;
    lbm_for_q([&](auto qc) {
      constexpr int q = LBM_QC(qc);
      constexpr int cx = LbmSet::c[q][0], cy = LbmSet::c[q][1], cz = LbmSet::c[q][2];
      // destination (x + cx, y + cy, z + cz) inside the lattice?
      const bool dst_in = (cx < 0 ? !x0 : (cx > 0 ? !xl : true)) && (cy < 0 ? !y0 : (cy > 0 ? !yl : true)) &&
                          (cz < 0 ? !z0 : (cz > 0 ? !zl : true));
      // destination on a driven face (x' = 0, x' = NX-1, y' = NY-1, z' = 0, z' = NZ-1)?
      const bool dst_driven = (cx < 0 ? x1 : (cx == 0 ? x0 : false)) || (cx > 0 ? xl1 : (cx == 0 ? xl : false)) ||
                              (cy > 0 ? yl1 : (cy == 0 ? yl : false)) ||
                              (cz < 0 ? z1 : (cz == 0 ? z0 : false)) || (cz > 0 ? zl1 : (cz == 0 ? zl : false));
      if (dst_in && !dst_driven) Fb[(unsigned int)q * un + (unsigned int)((int)cell + cx * sxy + cy * L.nz + cz)] = post[q];
      // this cell's own slot q
      if (driven) {
        Fb[(unsigned int)q * un + cell] = c_lbm.eq_force[q];     // stream.cs:25-35, after every push
      } else {
        // slot q is fed from (x - cx, y - cy, z - cz); outside the lattice nobody pushes into it: it keeps its value
        const bool src_in = (cx > 0 ? !x0 : (cx < 0 ? !xl : true)) && (cy > 0 ? !y0 : (cy < 0 ? !yl : true)) &&
                            (cz > 0 ? !z0 : (cz < 0 ? !zl : true));
        if (!src_in) Fb[(unsigned int)q * un + cell] = f[q];
      }
    });
  }
}

// move.cs:23-52: tracer particles drift with the trilinearly interpolated velocity
__global__ void k_lbm_advect(LbmDev L, int n, float4* __restrict__ pos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float P[4] = {pos[i].x, pos[i].y, pos[i].z, pos[i].w};
  int p[4], nn[4];
  float w[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { p[k] = (int)P[k]; nn[k] = p[k] + 1; w[k] = P[k] - (float)p[k]; }
  auto at = [&](int X, int Y, int Z) { return L.V[((size_t)X * L.ny + Y) * L.nz + Z]; };
  const float4 a000 = at(p[0], p[1], p[2]), a100 = at(nn[0], p[1], p[2]), a010 = at(p[0], nn[1], p[2]),
               a001 = at(p[0], p[1], nn[2]), a110 = at(nn[0], nn[1], p[2]), a101 = at(nn[0], p[1], nn[2]),
               a011 = at(p[0], nn[1], nn[2]), a111 = at(nn[0], nn[1], nn[2]);
  const float* v000 = &a000.x; const float* v100 = &a100.x; const float* v010 = &a010.x; const float* v001 = &a001.x;
  const float* v110 = &a110.x; const float* v101 = &a101.x; const float* v011 = &a011.x; const float* v111 = &a111.x;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float v00 = (1.0f - w[0]) * v000[k] + w[0] * v100[k];
    const float v01 = (1.0f - w[0]) * v001[k] + w[0] * v101[k];
    const float v10 = (1.0f - w[0]) * v010[k] + w[0] * v110[k];
    const float v11 = (1.0f - w[0]) * v011[k] + w[0] * v111[k];
    const float v0 = (1.0f - w[1]) * v00 + w[1] * v10;
    const float v1 = (1.0f - w[1]) * v01 + w[1] * v11;
    P[k] += (1.0f - w[2]) * v0 + w[2] * v1;
  }
  pos[i] = make_float4(P[0], P[1], P[2], P[3]);
}
