// umma_tile.cuh -- tile geometry and accumulator helpers shared by the tcgen05 MaxSim kernels
// (maxsim_umma.cu: one CTA per SM; maxsim_umma_pair.cu: CTA pairs, cta_group::2).
#pragma once
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace bms {

constexpr int kNumAccum = 4;                     // TMEM accumulators, kTileN fp32 columns each (4*128 = 512)
constexpr uint32_t kSubtileBytes = 128 * 128;    // 128 rows x 128 B (one swizzle-128B panel)
constexpr uint32_t kSmemLimit = 232448;          // 227 KB opt-in maximum per CTA
constexpr uint32_t kBarrierBytes = 2048;         // mbarriers + TMEM slot + the S4 exchange slab

template <int KIND>
struct Kind {
  // bf16: a 256 B row is two 128 B K-panels; s8: one.  Every tcgen05.mma step consumes 32 B of K per row.
  static constexpr int kPanels = KIND == 0 ? 2 : 1;
  static constexpr uint32_t kTileBytes = kPanels * kSubtileBytes;
  static constexpr int kPanelElems = KIND == 0 ? 64 : 128;  // TMA x-coordinate step per panel (elements)
  using Acc = typename std::conditional<KIND == 0, float, int>::type;
};

__device__ __forceinline__ float acc_from_bits(uint32_t v, float) { return __uint_as_float(v); }
__device__ __forceinline__ int acc_from_bits(uint32_t v, int) { return static_cast<int>(v); }
__device__ __forceinline__ float acc_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ int acc_max(int a, int b) { return max(a, b); }
__device__ __forceinline__ float acc_max3(float a, float b, float c) { return fmax3(a, b, c); }
__device__ __forceinline__ int acc_max3(int a, int b, int c) { return imax3(a, b, c); }

// max over the 32 columns a thread holds for one chunk (two independent chains for ILP)
template <typename Acc>
__device__ __forceinline__ Acc chunk_max(const uint32_t (&v)[32]) {
  Acc a = acc_from_bits(v[0], Acc{});
  Acc b = acc_from_bits(v[1], Acc{});
#pragma unroll
  for (int i = 2; i < 30; i += 4) {
    a = acc_max3(a, acc_from_bits(v[i], Acc{}), acc_from_bits(v[i + 1], Acc{}));
    b = acc_max3(b, acc_from_bits(v[i + 2], Acc{}), acc_from_bits(v[i + 3], Acc{}));
  }
  a = acc_max3(a, acc_from_bits(v[30], Acc{}), acc_from_bits(v[31], Acc{}));
  return acc_max(a, b);
}

}  // namespace bms
