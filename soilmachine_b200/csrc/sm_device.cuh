// sm_device.cuh -- device-side data layout and the accessor the sweep kernels hand to sm_core.cuh.
//
// HBM layout:
//   top[x*dimy+y]      Sec32 (32 B = one DRAM sector): the TOP section of every column, so
//                      height()/surface() (43-89 calls per particle-step in the reference) are one
//                      sector read and >92 % of add()/remove() calls are one sector write.
//   pool[slot]         Sec32: buried sections, chained by `below`; two free rings keyed by sweep
//                      parity so that slots freed in sweep s are reused from sweep s+1 on without
//                      any lock (frees append to ring[s&1], allocations pop ring[(s+1)&1]).
//   wfreq/wtrack/windfreq  f32[y*dimx+x]   (water.h:53,349; wind.h:50)
//   particle SoA       16-byte vectors per particle (pa/pb/pc)
//   bins               per-sweep cell-bin lists of live particles used for the id-ordered
//                      conflict detection (head[parity][bin] = tag<<32 | particle, next/key per
//                      particle); tags make clearing unnecessary.
#pragma once
#include <cuda_runtime.h>
#include "sm_core.cuh"

struct PoolRing {
  unsigned long long head;  // next entry to hand out (advanced only in sweeps of the other parity)
  unsigned long long tail;  // next entry to write   (advanced only in sweeps of this parity)
};

struct RunCtl {
  unsigned int barrier;         // monotone arrival counter of the grid barrier
  unsigned int alive_slot[3];   // live-particle totals, rotated per sweep
  unsigned int tag_base;        // first unused sweep tag
  unsigned int err;             // SM_ERR_* bits raised on device
  unsigned long long steps, sweeps, exit_oob, exit_evap, exit_stall, drops, alive;
  unsigned long long bump;      // pool high-water mark
  PoolRing ring[2];
};

struct DevCtx {
  Sec32* top;
  Sec32* pool;
  uint32_t* ringbuf[2];
  unsigned long long pool_cap;
  float* wfreq;
  float* wtrack;
  float* windfreq;
  const SoilDev* soils;
  int nsoils;
  int dimx, dimy, scale;
  RunCtl* ctl;
  // particle batch
  float4* pa;        // water: px,py,sx,sy        | wind: px,py,sx,sy
  double2* pb;       // water: volume,sediment    | wind: sediment,height
  uint2* pc;         // water: contains,-         | wind: contains, bits(sz)
  unsigned char* alive;
  unsigned int* done;            // tag of the last sweep this particle completed (0xFFFFFFFF = dead)
  unsigned long long* head[2];   // bin heads per sweep parity
  uint32_t* next[2];
  uint32_t* key[2];              // ipos packed x<<16|y
  int nbx, nby;                  // allocated bin grid (for the smallest bin edge)
};

// ---- memory helpers: everything mutable is read through L2 (the TU is compiled -dlcm=cg) ----------
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
  return *((const volatile unsigned int*)p);
}
__device__ __forceinline__ void st_volatile_u32(unsigned int* p, unsigned int v) {
  *((volatile unsigned int*)p) = v;
}

// Grid-wide barrier for a co-resident (cooperative) grid: one arrival per block on a monotone
// counter.  `epoch` is a per-thread copy of how many barriers this launch has passed.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& epoch) {
  __syncthreads();
  epoch++;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = epoch * gridDim.x;
    while ((int)(ld_volatile_u32(counter) - target) < 0) { }
    __threadfence();
  }
  __syncthreads();
}

// ---- accessor: direct global records ------------------------------------------------------------
struct DevAccess {
  const DevCtx& c;
  const SoilDev* s_soils;   // shared-memory copy of the soil table
  unsigned int phase;       // sweep parity for the pool rings
  __device__ __forceinline__ DevAccess(const DevCtx& ctx, const SoilDev* ss, unsigned int ph)
      : c(ctx), s_soils(ss), phase(ph & 1u) {}
  __device__ __forceinline__ int dimx() const { return c.dimx; }
  __device__ __forceinline__ int dimy() const { return c.dimy; }
  __device__ __forceinline__ int scale() const { return c.scale; }
  __device__ __forceinline__ SoilDev soil(uint32_t t) const { return s_soils[t]; }
  __device__ __forceinline__ Sec32* rec(int x, int y) { return &c.top[(size_t)x * c.dimy + y]; }
  __device__ __forceinline__ void begin(int, int) {}
  __device__ __forceinline__ void target(int, int) {}
  __device__ __forceinline__ void dirty(int, int) {}
  __device__ __forceinline__ Sec32 pool_load(uint32_t i) { return c.pool[i]; }
  __device__ __forceinline__ void pool_store(uint32_t i, const Sec32& r) { c.pool[i] = r; }
  __device__ uint32_t pool_alloc() {
    PoolRing* R = &c.ctl->ring[phase ^ 1u];
    unsigned long long t = R->tail;   // stable during this sweep
    unsigned long long h = *((volatile unsigned long long*)&R->head);
    while (h < t) {
      unsigned long long old = atomicCAS(&R->head, h, h + 1ull);
      if (old == h) return c.ringbuf[phase ^ 1u][h % c.pool_cap];
      h = old;
    }
    unsigned long long b = atomicAdd(&c.ctl->bump, 1ull);
    if (b < c.pool_cap) return (uint32_t)b;
    atomicOr(&c.ctl->err, 1u << 3);   // SM_ERR_POOL
    atomicAdd(&c.ctl->drops, 1ull);
    return SM_NIL;
  }
  __device__ void pool_free(uint32_t i) {
    PoolRing* R = &c.ctl->ring[phase];
    unsigned long long t = atomicAdd(&R->tail, 1ull);
    c.ringbuf[phase][t % c.pool_cap] = i;
  }
  __device__ __forceinline__ void track_add(int ind, double v) {     // water.h:348-351
    c.wtrack[ind] = (float)(c.wtrack[ind] + v);
  }
  __device__ __forceinline__ float water_frequency(int ind) { return c.wfreq[ind]; }
  __device__ __forceinline__ void wind_frequency_touch(int ind) {    // wind.h:49-52
    c.windfreq[ind] = (float)(0.5 * c.windfreq[ind] + 0.5f);
  }
};
