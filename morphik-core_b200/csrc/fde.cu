// fde.cu -- fixed-dimensional encodings (MUVERA FDE) and the dense candidate scan over them.
//
// Replaces stage 1-2 of FastMultiVectorStore.query_similar (core/vector_store/fast_multivector_store.py:521-532):
//   fde.generate_query_encoding / generate_document_encoding (C++ extension `fixed_dimensional_encoding`, configured at
//   :325-331 with dimension=128, num_repetitions=20, num_simhash_projections=5, projection_dimension=16, AMS_SKETCH)
//   + the Turbopuffer ANN query with distance_metric="cosine_distance" (:497,:526-532)
// by an exact device-side encoder and an exhaustive cosine scan of the [N, 10240] FDE matrix.
// The extension's sources are NOT in the reference snapshot (fde/ ships only pyproject.toml): the algorithm
// below restates the published MUVERA construction; the random matrices come from the host (fde.py) -- parity with the
// upstream RNG stream is unpinned and documented as such.
//
// Per repetition r:  partition(x) = gray-code index of the sign bits of x * G_r (G_r: [128, ksim] Gaussian),
//                    proj(x)[j]  = scale * sum_{i : idx_r[i] == j} sign_r[i] * x[i]        (AMS / count sketch, 128 -> proj)
//                    block[r][partition] += proj(x)      (queries: SUM;  documents: SUM / count = AVERAGE, empty -> 0)
// Deterministic: no atomics -- thread (partition, j) walks the rows in order, so the result is bit-identical to the
// oracle's loops (oracle/fde_oracle.c) in fp32.
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace bms {

constexpr int kFdeMaxSeg = 2048;  // rows whose partition ids are staged in shared memory at a time
constexpr int kFdeThreads = 512;  // = 32 partitions x 16 projection dims for the reference configuration

__device__ __forceinline__ float load_elem(const void* src, int src_dtype, int64_t row, int d) {
  if (src_dtype == B200MS_F32) return __ldg(static_cast<const float*>(src) + row * kDim + d);
  const uint16_t raw = __ldg(static_cast<const uint16_t*>(src) + row * kDim + d);
  return __uint_as_float(uint32_t(raw) << 16);
}

// grid = (n_items, reps); block = kFdeThreads
__global__ void __launch_bounds__(kFdeThreads)
fde_encode_kernel(const void* __restrict__ rows, int src_dtype, const int64_t* __restrict__ item_start,
                  const float* __restrict__ simhash /*[reps,128,ksim]*/, const int32_t* __restrict__ ams_index /*[reps,128]*/,
                  const float* __restrict__ ams_sign /*[reps,128]*/, int ksim, int proj, float scale, int is_document,
                  float* __restrict__ out /*[n_items, reps * 2^ksim * proj]*/) {
  __shared__ float s_g[kDim * 8];          // G_r, [128][ksim] (ksim <= 8)
  __shared__ int s_off[65];                // CSR of the AMS sketch: dims of bucket j are s_dims[s_off[j] .. s_off[j+1])
  __shared__ uint8_t s_dims[kDim];
  __shared__ float s_sgn[kDim];
  __shared__ uint8_t s_part[kFdeMaxSeg];
  const int item = blockIdx.x, rep = blockIdx.y, reps = gridDim.y;
  const int n_part = 1 << ksim;
  const int tid = threadIdx.x;
  for (int i = tid; i < kDim * ksim; i += blockDim.x) s_g[i] = simhash[int64_t(rep) * kDim * ksim + i];
  if (tid == 0) {  // counting sort of the 128 input dims by bucket (ascending dim inside a bucket)
    int cnt[64];
    for (int j = 0; j < proj; ++j) cnt[j] = 0;
    for (int i = 0; i < kDim; ++i) cnt[ams_index[rep * kDim + i]]++;
    int acc = 0;
    for (int j = 0; j < proj; ++j) {
      s_off[j] = acc;
      acc += cnt[j];
      cnt[j] = s_off[j];
    }
    s_off[proj] = acc;
    for (int i = 0; i < kDim; ++i) {
      const int j = ams_index[rep * kDim + i];
      s_dims[cnt[j]] = uint8_t(i);
      s_sgn[cnt[j]] = ams_sign[rep * kDim + i];
      cnt[j]++;
    }
  }
  __syncthreads();
  const int64_t r0 = item_start[item], r1 = item_start[item + 1];
  const int my_part = tid / proj, my_j = tid % proj;
  const bool worker = my_part < n_part;
  float acc = 0.f;
  int count = 0;
  for (int64_t seg = r0; seg < r1; seg += kFdeMaxSeg) {
    const int n = int(r1 - seg < kFdeMaxSeg ? r1 - seg : kFdeMaxSeg);
    // phase 1: SimHash partition of every row of the segment (one thread per row)
    for (int r = tid; r < n; r += blockDim.x) {
      float dot[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) dot[k] = 0.f;
      for (int d = 0; d < kDim; ++d) {
        const float x = load_elem(rows, src_dtype, seg + r, d);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < ksim) dot[k] = fmaf(x, s_g[d * ksim + k], dot[k]);
      }
      uint32_t code = 0;  // Gray-code append: idx = (idx << 1) + (bit ^ (idx & 1))
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < ksim) code = (code << 1) + (uint32_t(dot[k] > 0.f) ^ (code & 1u));
      s_part[r] = uint8_t(code);
    }
    __syncthreads();
    // phase 2: thread (partition, j) accumulates bucket j of the rows that fell into its partition, in row order
    if (worker) {
      const int e0 = s_off[my_j], e1 = s_off[my_j + 1];
      for (int r = 0; r < n; ++r) {
        if (s_part[r] != my_part) continue;
        count++;
        float v = 0.f;
        for (int e = e0; e < e1; ++e) v += s_sgn[e] * load_elem(rows, src_dtype, seg + r, s_dims[e]);
        acc += v;
      }
    }
    __syncthreads();
  }
  if (worker) {
    float v = acc * scale;
    if (is_document) v = count > 0 ? v / float(count) : 0.f;
    out[(int64_t(item) * reps + rep) * n_part * proj + my_part * proj + my_j] = v;
  }
}

// fp32 FDE rows -> bf16 rows of the FDE corpus + 1/||row|| (of the bf16-rounded row; 0 for an all-zero row)
__global__ void __launch_bounds__(256)
fde_finalize_kernel(const float* __restrict__ fde, int64_t n, int fde_dim, __nv_bfloat16* __restrict__ out_rows,
                    float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t warp_stride = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp_global; r < n; r += warp_stride) {
    float ss = 0.f;
    for (int d = lane; d < fde_dim; d += 32) {
      const __nv_bfloat16 b = __float2bfloat16_rn(fde[r * fde_dim + d]);
      out_rows[r * fde_dim + d] = b;
      const float f = __bfloat162float(b);
      ss = fmaf(f, f, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) inv_norm[r] = ss > 0.f ? rsqrtf(ss) : 0.f;
  }
}

// Exhaustive cosine scan: scores[q, p] = <q_fde[q], F[p]> * inv_norm[p]   (|q| is constant per query: same ranking as
// Turbopuffer's cosine_distance).  One warp per page row, 16-byte loads, up to kFdeQ queries resident in shared memory.
// HBM-bound: 2 * fde_dim bytes per page (20 KB for the reference configuration).
constexpr int kFdeQ = 4;
__global__ void __launch_bounds__(256)
fde_scan_kernel(const uint4* __restrict__ F, const float* __restrict__ inv_norm, int64_t n_pages, int fde_dim,
                const float* __restrict__ q_fde, int n_q, int q_base, float* __restrict__ scores, int64_t ld) {
  extern __shared__ float s_q[];  // [nq_here][fde_dim]
  const int nq_here = min(kFdeQ, n_q - q_base);
  for (int i = threadIdx.x; i < nq_here * fde_dim; i += blockDim.x) s_q[i] = q_fde[int64_t(q_base) * fde_dim + i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t warp_stride = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int vec_per_row = fde_dim / 8;  // uint4 = 8 bf16
  for (int64_t p = warp_global; p < n_pages; p += warp_stride) {
    float acc[kFdeQ];
#pragma unroll
    for (int q = 0; q < kFdeQ; ++q) acc[q] = 0.f;
    const uint4* row = F + p * vec_per_row;
    // 4 independent 16-byte loads in flight per lane (the scan is latency-bound otherwise: one warp per 20 KB row)
    for (int v0 = lane; v0 < vec_per_row; v0 += 128) {
      uint4 w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + 32 * u;
        w[u] = v < vec_per_row ? __ldg(row + v) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + 32 * u;
        if (v >= vec_per_row) break;
        float f[8];
        f[0] = __uint_as_float(w[u].x << 16); f[1] = __uint_as_float(w[u].x & 0xffff0000u);
        f[2] = __uint_as_float(w[u].y << 16); f[3] = __uint_as_float(w[u].y & 0xffff0000u);
        f[4] = __uint_as_float(w[u].z << 16); f[5] = __uint_as_float(w[u].z & 0xffff0000u);
        f[6] = __uint_as_float(w[u].w << 16); f[7] = __uint_as_float(w[u].w & 0xffff0000u);
#pragma unroll
        for (int q = 0; q < kFdeQ; ++q) {
          if (q < nq_here) {
            const float4* qq = reinterpret_cast<const float4*>(s_q + q * fde_dim + v * 8);
            const float4 q0 = qq[0], q1 = qq[1];
            acc[q] = fmaf(f[0], q0.x, acc[q]); acc[q] = fmaf(f[1], q0.y, acc[q]);
            acc[q] = fmaf(f[2], q0.z, acc[q]); acc[q] = fmaf(f[3], q0.w, acc[q]);
            acc[q] = fmaf(f[4], q1.x, acc[q]); acc[q] = fmaf(f[5], q1.y, acc[q]);
            acc[q] = fmaf(f[6], q1.z, acc[q]); acc[q] = fmaf(f[7], q1.w, acc[q]);
          }
        }
      }
    }
    const float inv = __ldg(inv_norm + p);
#pragma unroll
    for (int q = 0; q < kFdeQ; ++q) {
      if (q < nq_here) {
        const float s = warp_sum(acc[q]);
        if (lane == 0) scores[int64_t(q_base + q) * ld + p] = s * inv;
      }
    }
  }
}

int launch_fde_encode(b200ms_t* h, const void* rows, int src_dtype, const int64_t* item_start_dev, int n_items,
                      int is_document, float* out, cudaStream_t s) {
  if (n_items <= 0) return B200MS_OK;
  dim3 grid(n_items, h->fde_reps);
  fde_encode_kernel<<<grid, kFdeThreads, 0, s>>>(rows, src_dtype, item_start_dev, static_cast<const float*>(h->fde_simhash.p),
                                                static_cast<const int32_t*>(h->fde_ams_index.p),
                                                static_cast<const float*>(h->fde_ams_sign.p), h->fde_ksim, h->fde_proj,
                                                h->fde_scale, is_document, out);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "launch fde_encode");
}

int launch_fde_finalize(b200ms_t* h, const float* fde, int64_t n, void* out_rows, float* inv_norm, cudaStream_t s) {
  if (n <= 0) return B200MS_OK;
  int64_t blocks = (n + 7) / 8;
  if (blocks > int64_t(h->num_sms) * 16) blocks = int64_t(h->num_sms) * 16;
  fde_finalize_kernel<<<int(blocks), 256, 0, s>>>(fde, n, h->fde_dim, static_cast<__nv_bfloat16*>(out_rows), inv_norm);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "launch fde_finalize");
}

int launch_fde_scan(b200ms_t* h, const void* F, const float* inv_norm, int64_t n_pages, const float* q_fde, int n_q,
                    float* scores, int64_t ld, cudaStream_t s) {
  if (n_pages <= 0 || n_q <= 0) return B200MS_OK;
  const int nq_res = n_q < kFdeQ ? n_q : kFdeQ;  // queries resident per launch
  const size_t smem = size_t(nq_res) * h->fde_dim * sizeof(float);
  if (int e = check_cuda(h, cudaFuncSetAttribute(fde_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)),
                         "cudaFuncSetAttribute(fde_scan)"))
    return e;
  int per_sm = int((220 * 1024) / (smem + 1024));  // CTAs per SM by shared memory (40 KB per resident query)
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int64_t blocks = (n_pages + 7) / 8;
  if (blocks > int64_t(h->num_sms) * per_sm) blocks = int64_t(h->num_sms) * per_sm;
  for (int qb = 0; qb < n_q; qb += kFdeQ) {
    fde_scan_kernel<<<int(blocks), 256, smem, s>>>(static_cast<const uint4*>(F), inv_norm, n_pages, h->fde_dim, q_fde, n_q, qb,
                                                   scores, ld);
    h->launches++;
    if (int e = check_cuda(h, cudaGetLastError(), "launch fde_scan")) return e;
  }
  return B200MS_OK;
}

}  // namespace bms
