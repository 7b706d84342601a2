// wbc_wire.cuh — LCM wire codecs on the device: the byte formats either side of the control step.
//
//   trunk_state_t             lcm_types/trunk_state_t.lcm:1-50 (generated codec lcm_types/trunklcm/trunk_state_t.py:54-120),
//                             published by towr/trunk_mpc.cpp:19-68, consumed by planners/towr.py:37-48,92-148
//   robot_state_control_lcmt  lcm_types/robot_state_control_lcmt.lcm:1-7 (generated codec
//                             lcm_types/cheetahlcm/robot_state_control_lcmt.py:24-51), the use_lcm branch of
//                             controllers/basic_controller.py:79-87 (state in) and :307-317 (torques out)
//
// Wire format: 8-byte big-endian fingerprint, then the members in declaration order, big-endian, unpadded. Messages are
// packed back to back (549 / 204 bytes each). These are pure byte-moving kernels, HBM bound: a CTA stages a tile of
// messages in shared memory with coalesced 16-byte accesses and converts between the unaligned big-endian fields and
// the instance-major FP64 arrays of wbc.h with coalesced 8-byte accesses on the array side.
#pragma once
#include <stdint.h>
#include "wbc.h"

namespace wbcwire {

constexpr int TRUNK_B = WBC_LCM_TRUNK_STATE_BYTES;   // 549
constexpr int ROBOT_B = WBC_LCM_ROBOT_STATE_BYTES;   // 204
constexpr int TILE = 32;                             // messages per CTA tile (tile bytes are a multiple of 16)
constexpr int THREADS = 256;
// fingerprints = lcm-gen struct hash rotated left by one (trunk_state_t.py:122-128, robot_state_control_lcmt.py:53-59);
// stored as the big-endian wire bytes read as a little-endian u64
constexpr unsigned long long TRUNK_FP_BE = 0x7a078ad92c93a16dull;
constexpr unsigned long long ROBOT_FP_BE = 0x7c2811392475accfull;

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// 8 bytes at byte offset `off` of a 4-byte aligned shared buffer, as two words in memory order
__device__ __forceinline__ void load8(const uint32_t* sm, int off, uint32_t& first, uint32_t& second) {
  const int w = off >> 2, s = (off & 3) * 8;
  const uint32_t w0 = sm[w], w1 = sm[w + 1], w2 = sm[w + 2];     // w + 2 stays inside the tile (+ pad word)
  first = __funnelshift_r(w0, w1, s);
  second = __funnelshift_r(w1, w2, s);
}
__device__ __forceinline__ double load_be64(const uint32_t* sm, int off) {
  uint32_t a, b;
  load8(sm, off, a, b);
  return __hiloint2double((int)bswap32(a), (int)bswap32(b));
}
// 8 big-endian bytes of x at byte offset `off` (any alignment) of a 4-byte aligned shared buffer: the widest naturally
// aligned pieces instead of eight byte stores (2 stores when off is a multiple of 4, 3 when even, 4 when odd); neighbouring
// doubles share the boundary words, so the pieces never overlap another lane's bytes.
__device__ __forceinline__ void store_be64(unsigned char* sm, int off, double x) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  const uint32_t first = bswap32((uint32_t)(u >> 32)), second = bswap32((uint32_t)u);   // memory bytes 0-3 / 4-7, little-endian words
  unsigned char* p = sm + off;
  switch (off & 3) {
    case 0:
      *reinterpret_cast<uint32_t*>(p) = first; *reinterpret_cast<uint32_t*>(p + 4) = second;
      break;
    case 2:
      *reinterpret_cast<uint16_t*>(p) = (uint16_t)first;
      *reinterpret_cast<uint32_t*>(p + 2) = (first >> 16) | (second << 16);
      *reinterpret_cast<uint16_t*>(p + 6) = (uint16_t)(second >> 16);
      break;
    case 1:
      p[0] = (unsigned char)first;
      *reinterpret_cast<uint16_t*>(p + 1) = (uint16_t)(first >> 8);
      *reinterpret_cast<uint32_t*>(p + 3) = (first >> 24) | (second << 8);
      p[7] = (unsigned char)(second >> 24);
      break;
    default:
      p[0] = (unsigned char)first;
      *reinterpret_cast<uint32_t*>(p + 1) = (first >> 8) | (second << 24);
      *reinterpret_cast<uint16_t*>(p + 5) = (uint16_t)(second >> 8);
      p[7] = (unsigned char)(second >> 24);
      break;
  }
}

// Coalesced copy of `bytes` bytes between a 16-byte aligned global range and shared memory.
__device__ __forceinline__ void tile_load(unsigned char* sm, const unsigned char* g, int bytes) {
  const int nv = bytes >> 4;
  const uint4* g4 = reinterpret_cast<const uint4*>(g);
  uint4* s4 = reinterpret_cast<uint4*>(sm);
  for (int i = threadIdx.x; i < nv; i += blockDim.x) s4[i] = g4[i];
  for (int i = (nv << 4) + threadIdx.x; i < bytes; i += blockDim.x) sm[i] = g[i];
}
__device__ __forceinline__ void tile_store(unsigned char* g, const unsigned char* sm, int bytes) {
  const int nv = bytes >> 4;
  uint4* g4 = reinterpret_cast<uint4*>(g);
  const uint4* s4 = reinterpret_cast<const uint4*>(sm);
  for (int i = threadIdx.x; i < nv; i += blockDim.x) g4[i] = s4[i];
  for (int i = (nv << 4) + threadIdx.x; i < bytes; i += blockDim.x) g[i] = sm[i];
}

// ------------------------------------------------------------------ trunk_state_t -> traj / contact
// trunk_state_t.py:91-120 (_decode_one) for n messages; outputs of a message whose fingerprint does not match
// (the generated decoder raises ValueError, :85-86) are zero and its status is WBC_WIRE_BADFINGERPRINT.
__global__ void __launch_bounds__(THREADS) decode_trunk_kernel(const unsigned char* __restrict__ msgs, long long n,
                                                               double* __restrict__ timestamp, unsigned char* __restrict__ finished,
                                                               double* __restrict__ traj, unsigned char* __restrict__ contact,
                                                               double* __restrict__ fplan, int* __restrict__ status) {
  __shared__ __align__(16) unsigned char sm[TILE * TRUNK_B + 16];
  __shared__ int ok[TILE];
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(sm);
  for (long long t0 = (long long)blockIdx.x * TILE; t0 < n; t0 += (long long)gridDim.x * TILE) {
    const int cnt = (int)((n - t0) < TILE ? (n - t0) : TILE);
    __syncthreads();                                   // previous tile fully consumed
    tile_load(sm, msgs + t0 * TRUNK_B, cnt * TRUNK_B);
    __syncthreads();
    if (threadIdx.x < cnt) {
      const int m = threadIdx.x, base = m * TRUNK_B;
      uint32_t a, b;
      load8(sw, base, a, b);
      const bool good = (((unsigned long long)bswap32(a) << 32) | bswap32(b)) == TRUNK_FP_BE;
      ok[m] = good;
      if (status) status[t0 + m] = good ? 0 : WBC_WIRE_BADFINGERPRINT;
      if (timestamp) timestamp[t0 + m] = good ? load_be64(sw, base + 8) : 0.0;
      if (finished) finished[t0 + m] = good ? (sm[base + 16] != 0) : 0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * WBC_NTRAJ; e += THREADS) {
      const int m = e / WBC_NTRAJ, k = e - m * WBC_NTRAJ;
      traj[t0 * WBC_NTRAJ + e] = ok[m] ? load_be64(sw, m * TRUNK_B + 17 + 8 * k) : 0.0;
    }
    for (int e = threadIdx.x; e < cnt * 4; e += THREADS) {
      const int m = e >> 2, k = e & 3;
      contact[t0 * 4 + e] = ok[m] ? (sm[m * TRUNK_B + 449 + k] != 0) : 0;
    }
    if (fplan) {
      for (int e = threadIdx.x; e < cnt * 12; e += THREADS) {
        const int m = e / 12, k = e - m * 12;
        fplan[t0 * 12 + e] = ok[m] ? load_be64(sw, m * TRUNK_B + 453 + 8 * k) : 0.0;
      }
    }
  }
}

// trunk_state_t.py:54-88 (encode / _encode_one): what towr/trunk_mpc.cpp:19-68 publishes.
__global__ void __launch_bounds__(THREADS) encode_trunk_kernel(unsigned char* __restrict__ msgs, long long n,
                                                               const double* __restrict__ timestamp, const unsigned char* __restrict__ finished,
                                                               const double* __restrict__ traj, const unsigned char* __restrict__ contact,
                                                               const double* __restrict__ fplan) {
  __shared__ __align__(16) unsigned char sm[TILE * TRUNK_B + 16];
  for (long long t0 = (long long)blockIdx.x * TILE; t0 < n; t0 += (long long)gridDim.x * TILE) {
    const int cnt = (int)((n - t0) < TILE ? (n - t0) : TILE);
    __syncthreads();
    if (threadIdx.x < cnt) {
      const int m = threadIdx.x, base = m * TRUNK_B;
#pragma unroll
      for (int i = 0; i < 8; ++i) sm[base + i] = (unsigned char)(TRUNK_FP_BE >> (56 - 8 * i));
      store_be64(sm, base + 8, timestamp ? timestamp[t0 + m] : 0.0);
      sm[base + 16] = finished ? (finished[t0 + m] != 0) : 0;
    }
    for (int e = threadIdx.x; e < cnt * WBC_NTRAJ; e += THREADS) {
      const int m = e / WBC_NTRAJ, k = e - m * WBC_NTRAJ;
      store_be64(sm, m * TRUNK_B + 17 + 8 * k, traj[t0 * WBC_NTRAJ + e]);
    }
    for (int e = threadIdx.x; e < cnt * 4; e += THREADS) sm[(e >> 2) * TRUNK_B + 449 + (e & 3)] = contact[t0 * 4 + e] != 0;
    for (int e = threadIdx.x; e < cnt * 12; e += THREADS) {
      const int m = e / 12, k = e - m * 12;
      store_be64(sm, m * TRUNK_B + 453 + 8 * k, fplan ? fplan[t0 * 12 + e] : 0.0);
    }
    __syncthreads();
    tile_store(msgs + t0 * TRUNK_B, sm, cnt * TRUNK_B);
  }
}

// ------------------------------------------------------------------ robot_state_control_lcmt
constexpr int RTILE = 64;                              // 64 * 204 = 13056 bytes, a multiple of 16
constexpr int RWORDS = ROBOT_B / 4;                    // 51 words per message: 2 fingerprint + 19 q + 18 v + 12 tau

// robot_state_control_lcmt.py:41-47 (_decode_one) + basic_controller.py:85-87: float32 wire values widened to FP64.
__global__ void __launch_bounds__(THREADS) decode_robot_kernel(const unsigned char* __restrict__ msgs, long long n,
                                                               double* __restrict__ q, double* __restrict__ v, double* __restrict__ tau,
                                                               int* __restrict__ status) {
  __shared__ __align__(16) uint32_t sw[RTILE * RWORDS];
  for (long long t0 = (long long)blockIdx.x * RTILE; t0 < n; t0 += (long long)gridDim.x * RTILE) {
    const int cnt = (int)((n - t0) < RTILE ? (n - t0) : RTILE);
    __syncthreads();
    tile_load(reinterpret_cast<unsigned char*>(sw), msgs + t0 * ROBOT_B, cnt * ROBOT_B);
    __syncthreads();
    if (status && threadIdx.x < cnt) {
      const uint32_t* w = sw + threadIdx.x * RWORDS;
      const bool good = (((unsigned long long)bswap32(w[0]) << 32) | bswap32(w[1])) == ROBOT_FP_BE;
      status[t0 + threadIdx.x] = good ? 0 : WBC_WIRE_BADFINGERPRINT;
    }
    auto val = [&](int m, int j) {
      const uint32_t* w = sw + m * RWORDS;
      const bool good = (((unsigned long long)bswap32(w[0]) << 32) | bswap32(w[1])) == ROBOT_FP_BE;
      return good ? (double)__uint_as_float(bswap32(w[2 + j])) : 0.0;
    };
    for (int e = threadIdx.x; e < cnt * WBC_NQ; e += THREADS) { const int m = e / WBC_NQ; q[t0 * WBC_NQ + e] = val(m, e - m * WBC_NQ); }
    for (int e = threadIdx.x; e < cnt * WBC_NV; e += THREADS) { const int m = e / WBC_NV; v[t0 * WBC_NV + e] = val(m, 19 + e - m * WBC_NV); }
    if (tau)
      for (int e = threadIdx.x; e < cnt * WBC_NU; e += THREADS) { const int m = e / WBC_NU; tau[t0 * WBC_NU + e] = val(m, 37 + e - m * WBC_NU); }
  }
}

// robot_state_control_lcmt.py:24-33 (encode): FP64 -> float32 round-to-nearest-even -> big endian, which is what
// struct.pack('>f') does; a finite value that rounds to infinity makes struct.pack raise OverflowError -> WBC_WIRE_OVERFLOW.
// tau_map (12 ints or NULL): message slot j carries tau[tau_map[j]] — the reference sends the torques in velocity order,
// msg.tau = (S.T @ u)[-12:] (basic_controller.py:309-314), while wbc.h tau is in actuator order.
__global__ void __launch_bounds__(THREADS) encode_robot_kernel(unsigned char* __restrict__ msgs, long long n,
                                                               const double* __restrict__ q, const double* __restrict__ v,
                                                               const double* __restrict__ tau, const int* __restrict__ tau_map,
                                                               int* __restrict__ status) {
  __shared__ __align__(16) uint32_t sw[RTILE * RWORDS];
  __shared__ int ovf[RTILE];
  __shared__ int tmap[WBC_NU];
  if (threadIdx.x < WBC_NU) tmap[threadIdx.x] = tau_map ? tau_map[threadIdx.x] : threadIdx.x;
  for (long long t0 = (long long)blockIdx.x * RTILE; t0 < n; t0 += (long long)gridDim.x * RTILE) {
    const int cnt = (int)((n - t0) < RTILE ? (n - t0) : RTILE);
    __syncthreads();
    if (threadIdx.x < cnt) {
      sw[threadIdx.x * RWORDS] = bswap32((uint32_t)(ROBOT_FP_BE >> 32));
      sw[threadIdx.x * RWORDS + 1] = bswap32((uint32_t)ROBOT_FP_BE);
      ovf[threadIdx.x] = 0;
    }
    __syncthreads();
    auto put = [&](int m, int j, double x) {
      const float y = __double2float_rn(x);
      if (isinf(y) && !isinf(x)) ovf[m] = 1;           // benign race: every writer stores the same value
      sw[m * RWORDS + 2 + j] = bswap32(__float_as_uint(y));
    };
    for (int e = threadIdx.x; e < cnt * WBC_NQ; e += THREADS) { const int m = e / WBC_NQ; put(m, e - m * WBC_NQ, q ? q[t0 * WBC_NQ + e] : 0.0); }
    for (int e = threadIdx.x; e < cnt * WBC_NV; e += THREADS) { const int m = e / WBC_NV; put(m, 19 + e - m * WBC_NV, v ? v[t0 * WBC_NV + e] : 0.0); }
    for (int e = threadIdx.x; e < cnt * WBC_NU; e += THREADS) {
      const int m = e / WBC_NU, j = e - m * WBC_NU;
      put(m, 37 + j, tau[(t0 + m) * WBC_NU + tmap[j]]);
    }
    __syncthreads();
    if (status && threadIdx.x < cnt) status[t0 + threadIdx.x] = ovf[threadIdx.x] ? WBC_WIRE_OVERFLOW : 0;
    tile_store(msgs + t0 * ROBOT_B, reinterpret_cast<const unsigned char*>(sw), cnt * ROBOT_B);
  }
}

}  // namespace wbcwire
