// sm_noise.cuh -- terrain initialisation noise, bit-identical to what the reference evaluates.
//
// The reference fills the map with SurfLayer::get (surface.h:95-99) = max(min, bias + scale *
// FastNoiseLite::GetNoise(x, y, z)) configured as OpenSimplex2 + FBm (surface.h:82-89) from the
// vendored FastNoiseLite v1.0.1 (MIT).  This is a from-scratch restatement of that published
// algorithm for the one configuration the reference uses (float coordinates, 3-D, rotation type
// None => TransformType3D_DefaultOpenSimplex2, weighted strength 0, noise seed 1337 - the seed is
// never set, FastNoiseLite.h:114-116):
//   GetNoise / TransformNoiseCoordinate   FastNoiseLite.h:321-340, 686-727
//   GenFractalFBm, CalculateFractalBounding                 :865-885, 473-485
//   SingleOpenSimplex2 (3-D), GradCoord, Hash, FastRound    :1053-1150, 542-553, 500-506, 453
// All arithmetic is fp32 + wrapping 32-bit integer; FMA contraction must be off.
#pragma once
#include <stdint.h>
#include "sm_core.cuh"

struct LayerDev {       // SurfLayer (surface.h:65-101)
  uint32_t type;
  float min, bias, scale;
  int octaves;          // SetFractalOctaves((int)octaves)
  float lacunarity, gain, frequency;
  float bounding;       // mFractalBounding
};

SM_HD float fnl_fractal_bounding(int octaves, float gain_in) {   // FastNoiseLite.h:473-485
  float gain = gain_in < 0 ? -gain_in : gain_in;
  float amp = gain;
  float ampFractal = 1.0f;
  for (int i = 1; i < octaves; i++) {
    ampFractal += amp;
    amp *= gain;
  }
  return 1 / ampFractal;
}

#define FNL_PRIME_X 501125321u
#define FNL_PRIME_Y 1136930381u
#define FNL_PRIME_Z 1720413743u

// Gradients3D (FastNoiseLite.h:2529-2546): the 12 cube-edge directions five times over, then four
// extra entries; generated instead of tabulated.  idx = 0..63.
SM_HD void fnl_grad3(int idx, float& gx, float& gy, float& gz) {
  int g = idx;
  if (idx >= 60) {                       // (1,1,0) (0,-1,1) (-1,1,0) (0,-1,-1)
    const int e = idx - 60;
    gx = (e == 0) ? 1.f : (e == 2 ? -1.f : 0.f);
    gy = (e == 0 || e == 2) ? 1.f : -1.f;
    gz = (e == 1) ? 1.f : (e == 3 ? -1.f : 0.f);
    return;
  }
  g = idx % 12;
  const int grp = g >> 2, s = g & 3;     // grp 0: (0,+-1,+-1)  1: (+-1,0,+-1)  2: (+-1,+-1,0)
  const float a = (s & 1) ? -1.f : 1.f;  // first varying component
  const float b = (s & 2) ? -1.f : 1.f;  // second varying component
  if (grp == 0) { gx = 0.f; gy = a; gz = b; }
  else if (grp == 1) { gx = a; gy = 0.f; gz = b; }
  else { gx = a; gy = b; gz = 0.f; }
}

SM_HD float fnl_grad_coord(uint32_t seed, uint32_t xp, uint32_t yp, uint32_t zp, float xd, float yd, float zd) {
  uint32_t hash = seed ^ xp ^ yp ^ zp;          // Hash(), :500-506 (wrapping multiply)
  hash *= 0x27d4eb2du;
  int32_t h = (int32_t)hash;
  h ^= h >> 15;                                  // arithmetic shift on int
  h &= 63 << 2;
  float xg, yg, zg;
  fnl_grad3(h >> 2, xg, yg, zg);
  return xd * xg + yd * yg + zd * zg;
}

SM_HD int fnl_fast_round(float f) { return f >= 0 ? (int)(f + 0.5f) : (int)(f - 0.5f); }

SM_HD float fnl_single_opensimplex2(uint32_t seed, float x, float y, float z) {   // :1053-1150
  int i = fnl_fast_round(x), j = fnl_fast_round(y), k = fnl_fast_round(z);
  float x0 = (float)(x - i), y0 = (float)(y - j), z0 = (float)(z - k);
  int xNSign = (int)(-1.0f - x0) | 1;
  int yNSign = (int)(-1.0f - y0) | 1;
  int zNSign = (int)(-1.0f - z0) | 1;
  float ax0 = xNSign * -x0, ay0 = yNSign * -y0, az0 = zNSign * -z0;
  uint32_t ip = (uint32_t)i * FNL_PRIME_X, jp = (uint32_t)j * FNL_PRIME_Y, kp = (uint32_t)k * FNL_PRIME_Z;
  float value = 0;
  float a = (0.6f - x0 * x0) - (y0 * y0 + z0 * z0);
  for (int l = 0;; l++) {
    if (a > 0) value += (a * a) * (a * a) * fnl_grad_coord(seed, ip, jp, kp, x0, y0, z0);
    float b = a + 1;
    uint32_t i1 = ip, j1 = jp, k1 = kp;
    float x1 = x0, y1 = y0, z1 = z0;
    if (ax0 >= ay0 && ax0 >= az0) {
      x1 += xNSign;
      b -= xNSign * 2 * x1;
      i1 -= (uint32_t)xNSign * FNL_PRIME_X;
    } else if (ay0 > ax0 && ay0 >= az0) {
      y1 += yNSign;
      b -= yNSign * 2 * y1;
      j1 -= (uint32_t)yNSign * FNL_PRIME_Y;
    } else {
      z1 += zNSign;
      b -= zNSign * 2 * z1;
      k1 -= (uint32_t)zNSign * FNL_PRIME_Z;
    }
    if (b > 0) value += (b * b) * (b * b) * fnl_grad_coord(seed, i1, j1, k1, x1, y1, z1);
    if (l == 1) break;
    ax0 = 0.5f - ax0; ay0 = 0.5f - ay0; az0 = 0.5f - az0;
    x0 = xNSign * ax0; y0 = yNSign * ay0; z0 = zNSign * az0;
    a += (0.75f - ax0) - (ay0 + az0);
    ip += (uint32_t)(xNSign >> 1) & FNL_PRIME_X;
    jp += (uint32_t)(yNSign >> 1) & FNL_PRIME_Y;
    kp += (uint32_t)(zNSign >> 1) & FNL_PRIME_Z;
    xNSign = -xNSign; yNSign = -yNSign; zNSign = -zNSign;
    seed = ~seed;
  }
  return value * 32.69428253173828125f;
}

// GetNoise(x,y,z) with FractalType_FBm  (:321-340, 686-727, 865-885)
SM_HD float fnl_get_noise(const LayerDev& L, float x, float y, float z) {
  x *= L.frequency; y *= L.frequency; z *= L.frequency;
  {
    const float R3 = (float)(2.0 / 3.0);
    float r = (x + y + z) * R3;
    x = r - x; y = r - y; z = r - z;
  }
  uint32_t seed = 1337u;
  float sum = 0;
  float amp = L.bounding;
  for (int o = 0; o < L.octaves; o++) {
    float noise = fnl_single_opensimplex2(seed++, x, y, z);
    sum += noise * amp;
    amp *= 1.0f + 0.0f * ((noise + 1) * 0.5f - 1.0f);   // Lerp(1, (noise+1)/2, mWeightedStrength = 0)
    x *= L.lacunarity; y *= L.lacunarity; z *= L.lacunarity;
    amp *= L.gain;
  }
  return sum;
}

// SurfLayer::get at cell (i,j) for layer slice zslice  (surface.h:95-99, layermap.h:191)
SM_HD float layer_value(const LayerDev& L, int i, int j, int zslice, int dimx, int dimy) {
  // vec3(i, j, Z%MAXSEED) / vec3(dim.x, dim.y, 1)
  const float px = (float)i / (float)dimx, py = (float)j / (float)dimy, pz = (float)zslice / 1.0f;
  float val = L.bias + L.scale * fnl_get_noise(L, px, py, pz);
  if (val < L.min) val = L.min;
  return val;
}

// Z % MAXSEED for layer l of nlayers  (layermap.h:180-185)
inline int layer_zslice(int seed, int l, int nlayers) {
  const int MAXSEED = 10000;
  const float f = (float)l / (float)nlayers;
  const int Z = seed + f * MAXSEED;
  return Z % MAXSEED;
}
