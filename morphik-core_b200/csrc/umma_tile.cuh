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
constexpr uint32_t kBarrierBytes = 2048;         // mbarriers + TMEM slot

// KIND 0: bf16 rows (256 B = two 128 B K-panels), kind::f16, fp32 accumulators
// KIND 1: int8 rows (128 B = one panel),           kind::i8,  int32 accumulators (exact)
// KIND 2: fp8 e4m3 rows (128 B = one panel),       kind::f8f6f4, fp32 accumulators
// Every tcgen05.mma step consumes 32 B of K per row.
template <int KIND>
struct Kind {
  static constexpr int kPanels = KIND == 0 ? 2 : 1;
  static constexpr uint32_t kTileBytes = kPanels * kSubtileBytes;
  static constexpr int kPanelElems = KIND == 0 ? 64 : 128;  // TMA x-coordinate step per panel (elements)
  static constexpr int kKSteps = KIND == 0 ? 8 : 4;         // tcgen05.mma K steps per 128-d row
  static constexpr int kMaxNM = KIND == 0 ? 4 : 8;          // query tiles one CTA can keep resident in shared memory
  using Acc = typename std::conditional<KIND == 1, int, float>::type;
};

// corpus dtype (include/b200ms.h) -> kernel KIND, or -1
__host__ __device__ constexpr int kind_of_dtype(int dtype) {
  return dtype == B200MS_BF16 ? 0 : dtype == B200MS_I8 ? 1 : dtype == B200MS_F8 ? 2 : -1;
}

__device__ __forceinline__ float acc_from_bits(uint32_t v, float) { return __uint_as_float(v); }
__device__ __forceinline__ int acc_from_bits(uint32_t v, int) { return static_cast<int>(v); }
__device__ __forceinline__ float acc_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ int acc_max(int a, int b) { return max(a, b); }

// max over the 32 columns a thread holds for one chunk (two independent chains for ILP)
__device__ __forceinline__ float chunk_max_f(const uint32_t (&v)[32]) {
  float a = __uint_as_float(v[0]);
  float b = __uint_as_float(v[1]);
#pragma unroll
  for (int i = 2; i < 30; i += 4) {
    a = fmax3(a, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
    b = fmax3(b, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
  }
  a = fmax3(a, __uint_as_float(v[30]), __uint_as_float(v[31]));
  return fmaxf(a, b);
}
__device__ __forceinline__ int chunk_max_i(const uint32_t (&v)[32]) {
  int a = int(v[0]), b = int(v[1]);
#pragma unroll
  for (int i = 2; i < 30; i += 4) {
    a = imax3(a, int(v[i]), int(v[i + 1]));
    b = imax3(b, int(v[i + 2]), int(v[i + 3]));
  }
  a = imax3(a, int(v[30]), int(v[31]));
  return max(a, b);
}
// (Round 2 tried the int32 accumulators on the FLOAT min/max datapath -- bit patterns of non-negative ints below 0x7f800000 order
// like floats, with an integer redo when the float maximum came back negative / NaN.  Exact, but it bought nothing in the batch
// regime (TMEM-drain-bound, 6.16 vs 6.24 ms) and cost 37 % for ONE query, where a single warp's dependent max chain is the
// critical path: 1.42 ms integer vs 1.95 ms float at 65 536 pages, profiles/r02/ab_epilogue_variants.jsonl.  Integer it is.)
template <typename Acc>
__device__ __forceinline__ Acc chunk_max(const uint32_t (&v)[32]) {
  if constexpr (std::is_same<Acc, float>::value) {
    return chunk_max_f(v);
  } else {
    return chunk_max_i(v);
  }
}

// zero_pad_compat (option "zero_pad_compat"): colpali_engine's score_multi_vector pads every batch of pages with zero rows
// up to the longest page of the batch and lets them take part in the max (transformers port processing_colpali.py:350-362),
// so a page shorter than its batch's longest scores sum_t max(true max_t, 0).  clamp_bits bit i = "page (or candidate slot) i
// is shorter than the longest of its batch"; NULL = clean MaxSim (the default).
template <typename Acc>
__device__ __forceinline__ Acc clamp_token_max(const uint32_t* __restrict__ clamp_bits, int idx, Acc rm) {
  if (clamp_bits != nullptr && ((__ldg(clamp_bits + (idx >> 5)) >> (idx & 31)) & 1u)) return acc_max(rm, Acc(0));
  return rm;
}

}  // namespace bms
