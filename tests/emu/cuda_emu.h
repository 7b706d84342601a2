// Lock-step warp emulator: lets csrc/wbc_device.cuh be compiled for the host so the per-warp
// algorithm can be debugged without a GPU. TEST INFRASTRUCTURE ONLY - never part of the product
// library. Each lane is an OS thread; every warp collective is a pair of barriers.
#pragma once
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

struct EmuWarp {
  pthread_barrier_t bar;
  uint64_t slots[32];
};
extern thread_local int emu_lane;
extern thread_local EmuWarp* emu_warp;

static inline void emu_sync() { pthread_barrier_wait(&emu_warp->bar); }
template <typename T> static inline T emu_exchange(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  emu_warp->slots[emu_lane] = raw;
  emu_sync();
  uint64_t got = emu_warp->slots[src & 31];
  emu_sync();
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  const int base = emu_lane & ~(width - 1);
  return emu_exchange(v, base | (src & (width - 1)));
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int delta, int width = 32) {
  const int pos = emu_lane & (width - 1);
  return emu_exchange(v, pos + delta < width ? emu_lane + delta : emu_lane);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int delta, int width = 32) {
  const int pos = emu_lane & (width - 1);
  return emu_exchange(v, pos - delta >= 0 ? emu_lane - delta : emu_lane);
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
  (void)width;
  return emu_exchange(v, emu_lane ^ mask);
}
static inline int __any_sync(unsigned, int pred) {
  int acc = 0;
  for (int l = 0; l < 32; ++l) acc |= emu_exchange(pred ? 1 : 0, l);
  return acc;
}
static inline int __all_sync(unsigned m, int pred) { return !__any_sync(m, !pred); }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  unsigned acc = 0;
  for (int l = 0; l < 32; ++l) { const unsigned o = emu_exchange(v, l); acc = o > acc ? o : acc; }
  return acc;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  unsigned acc = 0xffffffffu;
  for (int l = 0; l < 32; ++l) { const unsigned o = emu_exchange(v, l); acc = o < acc ? o : acc; }
  return acc;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned acc = 0;
  for (int l = 0; l < 32; ++l) acc |= (unsigned)(emu_exchange(pred ? 1 : 0, l) & 1) << l;
  return acc;
}
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __double2hiint(double v) { long long r; memcpy(&r, &v, 8); return (int)(r >> 32); }
static inline int __double2loint(double v) { long long r; memcpy(&r, &v, 8); return (int)(r & 0xffffffffll); }
static inline double __hiloint2double(int hi, int lo) { long long r = ((long long)hi << 32) | (unsigned)lo; double d; memcpy(&d, &r, 8); return d; }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_sync(); }
static inline long long __double_as_longlong(double v) { long long r; memcpy(&r, &v, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
#define WBC_DEV static inline
