/* oracle/lbm_oracle.c -- CPU restatement of the reference's LBM wind field.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED.  Upstream runs this as OpenGL compute shaders (source/include/lbmwind/shader/LBM/*.cs,
 * driven by source/include/lbmwind/lbmwind.h:75-236); there is no GL in this image and the reference holds no
 * test or golden vector for it, so nothing of the reference can be executed to pin this file.  It restates the
 * shaders' arithmetic one statement at a time with an explicit evaluation order (fp32, left to right, no
 * contraction: compile with -ffp-contract=off) - that order is a DEFINITION where GLSL leaves freedom (a GLSL
 * compiler may contract a*b+c and folds the constants at a precision of its choosing).  The CUDA kernels
 * (soilmachine_b200/csrc/sm_lbm.cuh) are compared with this file bit for bit.
 *
 * Two more definitions where one shader dispatch races with itself upstream (stream.cs:7-38): a cell's pushed
 * populations and the boundary rewrite of the SAME dispatch have no order in GLSL; here the push happens first
 * and the rewrite of the five driven faces wins, which is the evident intent (wet-node driving).  Populations
 * that would arrive from outside the domain are never written and keep their previous value, as upstream.
 *
 * Layout as upstream: cell index ind = (x*NY + y)*NZ + z, populations F[ind*Q + q] (array of structures).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define Q 19
static const float Wq[Q] = {                                   /* lbm.cs:56-61 */
    1.0f / 3.0f,
    1.0f / 18.0f, 1.0f / 18.0f, 1.0f / 18.0f, 1.0f / 18.0f, 1.0f / 18.0f, 1.0f / 18.0f,
    1.0f / 36.0f, 1.0f / 36.0f, 1.0f / 36.0f, 1.0f / 36.0f, 1.0f / 36.0f, 1.0f / 36.0f,
    1.0f / 36.0f, 1.0f / 36.0f, 1.0f / 36.0f, 1.0f / 36.0f, 1.0f / 36.0f, 1.0f / 36.0f};
static const int Cq[Q][3] = {                                  /* lbm.cs:63-73 */
    {0, 0, 0},
    {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1},
    {1, 1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, -1}, {0, 1, 1}, {0, -1, -1},
    {1, -1, 0}, {-1, 1, 0}, {1, 0, -1}, {-1, 0, 1}, {0, 1, -1}, {0, -1, 1}};
static const int CP[Q] = {0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15, 18, 17};   /* lbm.cs:75-80 */

typedef struct {
  int nx, ny, nz;
  float *F, *FPROP, *B, *RHO, *V;   /* V: 4 floats per cell */
  float force[3];
  float cs2, cs4;
} Lbm;
static Lbm L;

static float dot3c(const float* v, int q) {     /* dot(v, c[q]): x*x + y*y + z*z, left to right */
  return v[0] * (float)Cq[q][0] + v[1] * (float)Cq[q][1] + v[2] * (float)Cq[q][2];
}
static float equilibrium(int q, float rho, const float* v) {   /* lbm.cs:88-97 */
  const float d = dot3c(v, q);
  float eq = 0.0f;
  eq += Wq[q] * rho;
  eq += Wq[q] * rho * d * L.cs2;
  eq += Wq[q] * rho * (d * d) * 0.5f * L.cs4;
  eq -= Wq[q] * rho * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) * 0.5f * L.cs2;
  return eq;
}
static float get_rho(size_t ind) {              /* lbm.cs:101-108 */
  float rho = 0.0f;
  for (int q = 0; q < Q; q++) rho += L.F[ind * Q + q];
  return rho;
}
static void get_v(size_t ind, float* v) {       /* lbm.cs:112-119 */
  v[0] = v[1] = v[2] = 0.0f;
  for (int q = 0; q < Q; q++)
    for (int k = 0; k < 3; k++) v[k] += L.F[ind * Q + q] * (float)Cq[q][k];
}

void lbmo_free(void) {
  free(L.F); free(L.FPROP); free(L.B); free(L.RHO); free(L.V);
  memset(&L, 0, sizeof(L));
}
/* lbmwind.h:75-118: buffers, then init.cs with the boundary still all zero */
void lbmo_create(int nx, int ny, int nz) {
  lbmo_free();
  L.nx = nx; L.ny = ny; L.nz = nz;
  const size_t n = (size_t)nx * ny * nz;
  L.F = (float*)calloc(n * Q, 4); L.FPROP = (float*)calloc(n * Q, 4);
  L.B = (float*)calloc(n, 4); L.RHO = (float*)calloc(n, 4); L.V = (float*)calloc(n * 4, 4);
  /* lbm.cs:35 force = 0.05*vec3(-2, 0, 1); lbm.cs:82-84 cs, cs2, cs4 */
  L.force[0] = 0.05f * -2.0f; L.force[1] = 0.05f * 0.0f; L.force[2] = 0.05f * 1.0f;
  const float cs = 1.0f / sqrtf(3.0f);
  L.cs2 = 1.0f / cs / cs;
  L.cs4 = 1.0f / cs / cs / cs / cs;
}
void lbmo_set_boundary(const float* b) { memcpy(L.B, b, (size_t)L.nx * L.ny * L.nz * 4); }

void lbmo_init(void) {                          /* init.cs:7-24 */
  const size_t n = (size_t)L.nx * L.ny * L.nz;
  const float zero[3] = {0.0f, 0.0f, 0.0f};
  for (size_t ind = 0; ind < n; ind++) {
    for (int q = 0; q < Q; q++) {
      L.F[ind * Q + q] = equilibrium(q, 1.0f, L.force);
      if (L.B[ind] > 0) L.F[ind * Q + q] = equilibrium(q, 1.0f, zero);
    }
    L.RHO[ind] = get_rho(ind);
    float v[3];
    get_v(ind, v);
    for (int k = 0; k < 3; k++) L.V[ind * 4 + k] = v[k] / L.RHO[ind];
    L.V[ind * 4 + 3] = 1.0f;
  }
}

static void collide(void) {                     /* collide.cs:10-58 */
  const size_t n = (size_t)L.nx * L.ny * L.nz;
  const float tau = 0.56f, dt = 1.0f;
  const float omega_plus = 1.0f / tau;
  const float lambda = 0.25f;
  const float omega_minus = 1.0f / (lambda / (1.0f / omega_plus - 0.5f) + 0.5f);
  const float zero[3] = {0.0f, 0.0f, 0.0f};
  for (size_t ind = 0; ind < n; ind++) {
    const float rho = get_rho(ind);
    float v[3];
    get_v(ind, v);
    for (int k = 0; k < 3; k++) v[k] = v[k] / rho;
    for (int k = 0; k < 3; k++) v[k] += dt * 0.0001f * (float)Cq[4][k] / (2.0f * rho);   /* gravity, :21 */
    L.RHO[ind] = rho;
    for (int k = 0; k < 3; k++) L.V[ind * 4 + k] = v[k];
    L.V[ind * 4 + 3] = 0.0f;
    float ffeq[Q];
    for (int q = 0; q < Q; q++) ffeq[q] = equilibrium(q, rho, v);
    for (int q = 0; q < Q; q++) {
      const float f_plus = 0.5f * (L.F[ind * Q + q] + L.F[ind * Q + CP[q]]);
      const float f_minus = 0.5f * (L.F[ind * Q + q] - L.F[ind * Q + CP[q]]);
      const float feq_plus = 0.5f * (ffeq[q] + ffeq[CP[q]]);
      const float feq_minus = 0.5f * (ffeq[q] - ffeq[CP[q]]);
      L.FPROP[ind * Q + q] = L.F[ind * Q + q] - omega_plus * (f_plus - feq_plus) - omega_minus * (f_minus - feq_minus);
      if (L.B[ind] > 0.0f) L.FPROP[ind * Q + q] = equilibrium(q, 1.0f, zero);
    }
  }
}
static void stream(void) {                      /* stream.cs:7-38 */
  for (int x = 0; x < L.nx; x++) for (int y = 0; y < L.ny; y++) for (int z = 0; z < L.nz; z++) {
    const size_t ind = ((size_t)x * L.ny + y) * L.nz + z;
    for (int q = 0; q < Q; q++) {
      const int nx = x + Cq[q][0], ny = y + Cq[q][1], nz = z + Cq[q][2];
      if (nx < 0 || nx >= L.nx) continue;
      if (ny < 0 || ny >= L.ny) continue;
      if (nz < 0 || nz >= L.nz) continue;
      L.F[(((size_t)nx * L.ny + ny) * L.nz + nz) * Q + q] = L.FPROP[ind * Q + q];
    }
  }
  /* the driven faces, after every push of this step (see the header) */
  for (int x = 0; x < L.nx; x++) for (int y = 0; y < L.ny; y++) for (int z = 0; z < L.nz; z++) {
    if (y == L.ny - 1 || x == 0 || x == L.nx - 1 || z == 0 || z == L.nz - 1) {
      const size_t ind = ((size_t)x * L.ny + y) * L.nz + z;
      for (int q = 0; q < Q; q++) L.F[ind * Q + q] = equilibrium(q, 1.0f, L.force);
    }
  }
}
void lbmo_step(int n) {                         /* lbmwind.h:170-187: collide, then stream */
  for (int i = 0; i < n; i++) { collide(); stream(); }
}
void lbmo_get(float* f, float* rho, float* v4) {
  const size_t n = (size_t)L.nx * L.ny * L.nz;
  if (f) memcpy(f, L.F, n * Q * 4);
  if (rho) memcpy(rho, L.RHO, n * 4);
  if (v4) memcpy(v4, L.V, n * 16);
}
/* move.cs:23-52: advect tracer particles through V (trilinear); pos4 = n x (x, y, z, w) */
void lbmo_advect(int n, float* pos4) {
  for (int i = 0; i < n; i++) {
    float* P = pos4 + 4 * (size_t)i;
    int p[4], nn[4];
    float w[4];
    for (int k = 0; k < 4; k++) { p[k] = (int)P[k]; nn[k] = p[k] + 1; w[k] = P[k] - (float)p[k]; }
#define VAT(X, Y, Z) (&L.V[((((size_t)(X)) * L.ny + (Y)) * L.nz + (Z)) * 4])
    const float *v000 = VAT(p[0], p[1], p[2]), *v100 = VAT(nn[0], p[1], p[2]), *v010 = VAT(p[0], nn[1], p[2]),
                *v001 = VAT(p[0], p[1], nn[2]), *v110 = VAT(nn[0], nn[1], p[2]), *v101 = VAT(nn[0], p[1], nn[2]),
                *v011 = VAT(p[0], nn[1], nn[2]), *v111 = VAT(nn[0], nn[1], nn[2]);
#undef VAT
    for (int k = 0; k < 4; k++) {
      const float v00 = (1.0f - w[0]) * v000[k] + w[0] * v100[k];
      const float v01 = (1.0f - w[0]) * v001[k] + w[0] * v101[k];
      const float v10 = (1.0f - w[0]) * v010[k] + w[0] * v110[k];
      const float v11 = (1.0f - w[0]) * v011[k] + w[0] * v111[k];
      const float v0 = (1.0f - w[1]) * v00 + w[1] * v10;
      const float v1 = (1.0f - w[1]) * v01 + w[1] * v11;
      P[k] += (1.0f - w[2]) * v0 + w[2] * v1;
    }
  }
}
