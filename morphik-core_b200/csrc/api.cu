// api.cu -- the extern "C" surface declared in include/b200ms.h: handle management, corpus attachment
// (chunk table, work-unit plan, TMA descriptor), query packing and the fused search entry points.
// Host-side C++ only orchestrates; all arithmetic on embeddings happens in the CUDA kernels.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace bms {

static std::mutex g_err_mu;
static std::string g_err;

int set_error(b200ms_t* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_err = msg;
  return code;
}

int check_cuda(b200ms_t* h, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return B200MS_OK;
  return set_error(h, B200MS_ECUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

int reserve(b200ms_t* h, DeviceBuf& b, size_t bytes) {
  if (bytes <= b.cap) return B200MS_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 4;
  want = (want + 1023) & ~size_t(1023);
  if (cudaMalloc(&b.p, want) != cudaSuccess) {
    cudaGetLastError();
    want = (bytes + 1023) & ~size_t(1023);
    if (cudaError_t e = cudaMalloc(&b.p, want); e != cudaSuccess) {
      b.p = nullptr;
      cudaGetLastError();
      return set_error(h, B200MS_ENOMEM, "cudaMalloc of " + std::to_string(want) + " scratch bytes failed");
    }
  }
  b.cap = want;
  return B200MS_OK;
}

int upload(b200ms_t* h, DeviceBuf& b, const void* src, size_t bytes, cudaStream_t s) {
  if (int e = reserve(h, b, bytes ? bytes : 16)) return e;
  if (bytes == 0) return B200MS_OK;
  return check_cuda(h, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s), "upload metadata");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn(b200ms_t* h) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !p) {
      cudaGetLastError();
      set_error(h, B200MS_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [n_rows, 128] row-major view; box = one 128-byte K panel x box_rows rows, 128-byte swizzle (what the UMMA
// K-major SWIZZLE_128B descriptor expects).  Out-of-range rows read as zero.
int make_tmap_rows(b200ms_t* h, CUtensorMap* out, const void* base, int dtype, int64_t n_rows, int box_rows) {
  EncodeTiledFn fn = encode_fn(h);
  if (!fn) return B200MS_ECUDA;
  const bool bf16 = dtype == B200MS_BF16;
  const cuuint64_t dims[2] = {cuuint64_t(kDim), cuuint64_t(n_rows > 0 ? n_rows : 1)};
  const cuuint64_t strides[1] = {cuuint64_t(bf16 ? 256 : 128)};
  const cuuint32_t box[2] = {cuuint32_t(bf16 ? 64 : 128), cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(h, B200MS_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
  return B200MS_OK;
}

static bool corpus_dtype_ok(int d) { return d == B200MS_BF16 || d == B200MS_I8 || d == B200MS_B1; }
static bool src_dtype_ok(int d) { return d == B200MS_F32 || d == B200MS_BF16; }

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

}  // namespace bms

using namespace bms;

#define B200MS_API extern "C" __attribute__((visibility("default")))

B200MS_API int b200ms_version(void) { return B200MS_VERSION; }

B200MS_API int b200ms_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

B200MS_API const char* b200ms_last_error(const b200ms_t* h) {
  if (h) return h->err.c_str();
  std::lock_guard<std::mutex> lk(g_err_mu);
  return g_err.c_str();
}

B200MS_API int b200ms_create(int device, b200ms_t** out) {
  if (!out) return set_error(nullptr, B200MS_EINVAL, "create: out is NULL");
  *out = nullptr;
  int n = 0;
  if (cudaError_t e = cudaGetDeviceCount(&n); e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_error(nullptr, B200MS_ECUDA, "create: no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0 || device >= n) return set_error(nullptr, B200MS_EINVAL, "create: bad device index");
  DeviceGuard g(device);
  cudaDeviceProp prop;
  if (int e = check_cuda(nullptr, cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return e;
  if (prop.major != 10) {
    return set_error(nullptr, B200MS_ECUDA, std::string("create: device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                                ", this library is built for sm_100a (B200) only");
  }
  b200ms_t* h = new b200ms_t();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("B200MS_A_IN_TMEM")) h->a_in_tmem = atoi(e) != 0;
  if (const char* e = getenv("B200MS_B1_TENSOR")) h->b1_tensor = atoi(e);
  if (const char* e = getenv("B200MS_SPLIT4")) h->split4 = atoi(e);
  if (const char* e = getenv("B200MS_EPI_W4")) h->epi_w4 = atoi(e) != 0;
  if (const char* e = getenv("B200MS_PAIR_CTA")) h->pair_cta = atoi(e);
  if (const char* e = getenv("B200MS_UNIT_ROWS")) { if (atoll(e) > 0) h->unit_rows = atoll(e); }
  if (int e = check_cuda(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate")) {
    delete h;
    return e;
  }
  for (int i = 0; i < b200ms_t::kEvRing; ++i) {
    cudaEventCreate(&h->ev0[i]);
    cudaEventCreate(&h->ev1[i]);
  }
  *out = h;
  return B200MS_OK;
}

B200MS_API int b200ms_destroy(b200ms_t* h) {
  if (!h) return B200MS_OK;
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();
  DeviceBuf* bufs[] = {&h->chunk_page, &h->unit_start, &h->page_start, &h->meta_a, &h->meta_b, &h->meta_c, &h->q_raw,
                       &h->q_packed, &h->scores, &h->mask, &h->out_s, &h->out_i, &h->out_c, &h->cand_start, &h->cand_end,
                       &h->cand_mask, &h->topk_keys, &h->topk_ids, &h->b1_q_i8, &h->b1_tok_const, &h->fde_simhash, &h->fde_ams_index, &h->fde_ams_sign, &h->fde_tmp};
  for (DeviceBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  for (int i = 0; i < b200ms_t::kEvRing; ++i) {
    if (h->ev0[i]) cudaEventDestroy(h->ev0[i]);
    if (h->ev1[i]) cudaEventDestroy(h->ev1[i]);
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return B200MS_OK;
}

B200MS_API int64_t b200ms_padded_len(int64_t len) { return len <= 0 ? 0 : (len + kGroup - 1) / kGroup * kGroup; }

B200MS_API int64_t b200ms_padded_rows(const int32_t* page_lens, int64_t n_pages) {
  int64_t t = 0;
  for (int64_t i = 0; i < n_pages; ++i) t += b200ms_padded_len(page_lens[i]);
  return t;
}

B200MS_API int64_t b200ms_row_bytes(int dtype) {
  switch (dtype) {
    case B200MS_F32: return 512;
    case B200MS_BF16: return 256;
    case B200MS_I8: return 128;
    case B200MS_B1: return 16;
    default: return 0;
  }
}

B200MS_API int64_t b200ms_query_groups(const int32_t* q_lens, int n_q) {
  int64_t g = 0;
  for (int i = 0; i < n_q; ++i) g += (int64_t(q_lens[i] > 0 ? q_lens[i] : 0) + kGroup - 1) / kGroup;
  return g;
}

B200MS_API int64_t b200ms_launch_count(const b200ms_t* h) { return h ? h->launches : 0; }

B200MS_API int b200ms_set_tuning(b200ms_t* h, int64_t unit_rows, int max_ctas) {
  if (!h) return B200MS_EINVAL;
  if (unit_rows > 0) h->unit_rows = unit_rows;
  if (max_ctas >= 0) h->max_ctas = max_ctas;
  return B200MS_OK;
}

B200MS_API int b200ms_set_option(b200ms_t* h, const char* name, int64_t value) {
  if (!h || !name) return B200MS_EINVAL;
  const std::string n(name);
  if (n == "a_in_tmem") {
    h->a_in_tmem = value != 0;
  } else if (n == "epi_w4") {
    h->epi_w4 = value != 0;
  } else if (n == "pair_cta") {
    h->pair_cta = int(value);
  } else if (n == "split4" && value >= 0 && value <= 2) {
    h->split4 = int(value);
  } else if (n == "b1_tensor" && value >= 0 && value <= 2) {
    h->b1_tensor = int(value);
  } else if (n == "unit_rows" && value > 0) {
    h->unit_rows = value;
  } else if (n == "max_ctas" && value >= 0) {
    h->max_ctas = int(value);
  } else {
    return set_error(h, B200MS_EINVAL, "set_option: unknown option or bad value: " + n);
  }
  return B200MS_OK;
}

B200MS_API int64_t b200ms_corpus_pages(const b200ms_t* h) { return h ? h->corpus.n_pages : 0; }
B200MS_API int64_t b200ms_corpus_rows(const b200ms_t* h) { return h ? h->corpus.n_rows : 0; }

// ------------------------------------------------------------------------------------------------ quantise / pack
static int pack_items(b200ms_t* h, const void* src, int src_dtype, const int32_t* lens, int64_t n, bool queries,
                      void* dst, int dst_dtype, int64_t dst_rows_total, float i8_scale, cudaStream_t s,
                      std::vector<int64_t>* dst_start_out) {
  std::vector<int64_t> ss(size_t(n) + 1), ds(size_t(n) + 1);
  ss[0] = ds[0] = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t len = lens[i] > 0 ? lens[i] : 0;
    ss[i + 1] = ss[i] + len;
    ds[i + 1] = ds[i] + b200ms_padded_len(len);
  }
  const int64_t dst_rows = dst_rows_total >= 0 ? dst_rows_total : ds[n];
  if (int e = upload(h, h->meta_a, ss.data(), ss.size() * 8, s)) return e;
  if (int e = upload(h, h->meta_b, ds.data(), ds.size() * 8, s)) return e;
  if (int e = launch_pack_rows(h, src, src_dtype, static_cast<const int64_t*>(h->meta_a.p),
                               static_cast<const int64_t*>(h->meta_b.p), n, dst_rows, queries ? 1 : 0, dst, dst_dtype,
                               i8_scale, s))
    return e;
  // ss/ds are pageable host memory: cudaMemcpyAsync staged them before returning, so they may die here
  if (dst_start_out) *dst_start_out = std::move(ds);
  return B200MS_OK;
}

B200MS_API int b200ms_sign_pack(b200ms_t* h, const void* x, int src_dtype, int64_t rows, uint8_t* out, void* stream) {
  if (!h) return B200MS_EINVAL;
  if (!src_dtype_ok(src_dtype) || rows < 0 || (rows > 0 && (!x || !out)))
    return set_error(h, B200MS_EINVAL, "sign_pack: bad arguments");
  if (rows == 0) return B200MS_OK;
  DeviceGuard g(h->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t st[2] = {0, rows};
  if (int e = upload(h, h->meta_a, st, sizeof(st), s)) return e;
  const int64_t* d = static_cast<const int64_t*>(h->meta_a.p);
  return launch_pack_rows(h, x, src_dtype, d, d, 1, rows, 0, out, B200MS_B1, 1.f, s);
}

B200MS_API int b200ms_hamming_batch(b200ms_t* h, const uint8_t* q_bits, const uint8_t* cand_bits, int64_t n, uint32_t* out,
                                    void* stream) {
  if (!h) return B200MS_EINVAL;
  if (n < 0 || (n > 0 && (!q_bits || !cand_bits || !out)) || (reinterpret_cast<uintptr_t>(q_bits) & 15) ||
      (reinterpret_cast<uintptr_t>(cand_bits) & 15))
    return set_error(h, B200MS_EINVAL, "hamming_batch: bad arguments (16-byte aligned packed rows)");
  DeviceGuard g(h->device);
  return launch_hamming_batch(h, q_bits, cand_bits, n, out, static_cast<cudaStream_t>(stream));
}

B200MS_API int b200ms_pack_pages(b200ms_t* h, const void* src, int src_dtype, const int32_t* page_lens, int64_t n_pages,
                                 void* dst, int dst_dtype, float i8_scale, void* stream) {
  if (!h) return B200MS_EINVAL;
  if (!src_dtype_ok(src_dtype) || !corpus_dtype_ok(dst_dtype) || n_pages < 0 || (n_pages > 0 && (!page_lens || !dst)))
    return set_error(h, B200MS_EINVAL, "pack_pages: bad arguments");
  if (n_pages == 0) return B200MS_OK;
  DeviceGuard g(h->device);
  return pack_items(h, src, src_dtype, page_lens, n_pages, false, dst, dst_dtype, -1, i8_scale,
                    static_cast<cudaStream_t>(stream), nullptr);
}

B200MS_API int b200ms_pack_queries(b200ms_t* h, const void* q, int src_dtype, const int32_t* q_lens, int n_q,
                                   void* q_packed, int dst_dtype, float i8_scale, int32_t* group_offsets_out,
                                   int* n_groups_out, void* stream) {
  if (!h) return B200MS_EINVAL;
  if (!src_dtype_ok(src_dtype) || !corpus_dtype_ok(dst_dtype) || n_q < 0 || (n_q > 0 && (!q_lens || !q_packed)))
    return set_error(h, B200MS_EINVAL, "pack_queries: bad arguments");
  DeviceGuard g(h->device);
  const int64_t groups = b200ms_query_groups(q_lens, n_q);
  const int64_t groups_padded = (groups + 3) & ~int64_t(3);
  std::vector<int64_t> ds;
  if (groups_padded > 0) {
    if (int e = pack_items(h, q, src_dtype, q_lens, n_q, true, q_packed, dst_dtype, groups_padded * kGroup, i8_scale,
                           static_cast<cudaStream_t>(stream), &ds))
      return e;
  } else {
    ds.assign(size_t(n_q) + 1, 0);
  }
  if (group_offsets_out)
    for (int i = 0; i <= n_q; ++i) group_offsets_out[i] = int32_t(ds[i] / kGroup);
  if (n_groups_out) *n_groups_out = int(groups);
  return B200MS_OK;
}

// ------------------------------------------------------------------------------------------------ corpus
B200MS_API int b200ms_set_corpus(b200ms_t* h, const void* rows, int dtype, const int32_t* page_lens, int64_t n_pages) {
  if (!h) return B200MS_EINVAL;
  if (!corpus_dtype_ok(dtype) || n_pages < 0 || (n_pages > 0 && (!rows || !page_lens)))
    return set_error(h, B200MS_EINVAL, "set_corpus: bad arguments");
  if (n_pages > 0 && (reinterpret_cast<uintptr_t>(rows) & 1023))
    return set_error(h, B200MS_EINVAL, "set_corpus: rows must be 1024-byte aligned");
  DeviceGuard g(h->device);
  cudaStream_t s = h->stream;
  Corpus c;
  c.rows = rows;
  c.dtype = dtype;
  c.n_pages = n_pages;
  std::vector<int64_t> ps(size_t(n_pages) + 1);
  ps[0] = 0;
  for (int64_t i = 0; i < n_pages; ++i) {
    ps[i + 1] = ps[i] + b200ms_padded_len(page_lens[i]);
    if (page_lens[i] <= 0) c.has_empty = true;
  }
  c.n_rows = ps[n_pages];
  c.n_chunks = c.n_rows / kGroup;
  if (c.n_chunks >= (int64_t(1) << 31)) return set_error(h, B200MS_EINVAL, "set_corpus: more than 2^31 chunks");
  // work units: runs of whole pages of about unit_rows rows; unit_start in chunks
  // (small corpora: shrink the units so that every SM still gets about four of them -- a 100-page corpus as 25 units
  // of 4096 rows would leave 123 SMs idle)
  int64_t unit_rows = c.n_rows / (4 * int64_t(h->num_sms > 0 ? h->num_sms : 1));
  if (unit_rows < kTileN) unit_rows = kTileN;
  if (unit_rows > h->unit_rows) unit_rows = h->unit_rows;
  std::vector<int32_t> us;
  us.push_back(0);
  int64_t acc = 0;
  for (int64_t i = 0; i < n_pages; ++i) {
    acc += ps[i + 1] - ps[i];
    if (acc >= unit_rows) {
      us.push_back(int32_t(ps[i + 1] / kGroup));
      acc = 0;
    }
  }
  if (us.back() != int32_t(c.n_chunks)) us.push_back(int32_t(c.n_chunks));
  c.n_units = int(us.size()) - 1;
  if (int e = upload(h, h->page_start, ps.data(), ps.size() * 8, s)) return e;
  if (int e = upload(h, h->unit_start, us.data(), us.size() * 4, s)) return e;
  if (int e = reserve(h, h->chunk_page, size_t(c.n_chunks > 0 ? c.n_chunks : 1) * 4)) return e;
  if (int e = launch_chunk_page(h, static_cast<const int64_t*>(h->page_start.p), n_pages,
                                static_cast<int32_t*>(h->chunk_page.p), s))
    return e;
  if (dtype != B200MS_B1 && c.n_rows > 0) {
    if (int e = make_tmap_rows(h, &c.tmap, rows, dtype, c.n_rows, kTileN)) return e;
    c.has_tmap = true;
  }
  if (int e = check_cuda(h, cudaStreamSynchronize(s), "set_corpus: stream sync")) return e;
  h->corpus = c;
  return B200MS_OK;
}

// ------------------------------------------------------------------------------------------------ hot path
// cand_ids == NULL: scan the whole corpus, scores indexed by page id.  Otherwise: score only the n_cand candidate pages
// (device array of page ids, -1 = unused slot), scores indexed by candidate slot.
static int score_impl(b200ms_t* h, const void* q_packed, int n_groups, const int32_t* q_lens,
                      const int32_t* group_offsets, int n_q, void* group_scores, int64_t ld, const int64_t* cand_ids,
                      int n_cand, cudaStream_t s) {
  const Corpus& c = h->corpus;
  if (c.dtype < 0) return set_error(h, B200MS_ESTATE, "score: no corpus attached (call b200ms_set_corpus first)");
  const int64_t n_items = cand_ids ? n_cand : c.n_pages;
  if (n_groups < 0 || ld < n_items || n_cand < 0 || (n_groups > 0 && (!q_packed || !group_scores)))
    return set_error(h, B200MS_EINVAL, "score: bad arguments");
  if (n_groups == 0 || n_items == 0 || c.n_pages == 0) return B200MS_OK;
  const int slot = int(h->ev_count % b200ms_t::kEvRing);
  const int n_groups_padded = (n_groups + 3) & ~3;
  if (c.dtype == B200MS_B1) {
    if (!q_lens || !group_offsets || n_q <= 0) return set_error(h, B200MS_EINVAL, "score: B1 needs q_lens and group_offsets");
    std::vector<int32_t> ntok(size_t(n_groups_padded), 0);
    for (int q = 0; q < n_q; ++q) {
      int left = q_lens[q];
      for (int g = group_offsets[q]; g < group_offsets[q + 1] && g < n_groups; ++g) {
        ntok[g] = left > kGroup ? kGroup : left;
        left -= ntok[g];
      }
    }
    if (int e = upload(h, h->meta_c, ntok.data(), ntok.size() * 4, s)) return e;
    // b1_tensor: 0 = POPC kernel, 1 = tcgen05 kernel, 2 (default) = measured crossover: one 32-token group is faster on the
    // POPC pipe (1.59 vs 2.49 ms / 65536 pages), two or more groups on the tensor cores (a 128-token tile costs the same
    // as one token there).  Rerank of a candidate list stays on the POPC kernel.
    const bool tensor_path = !cand_ids && (h->b1_tensor == 1 || (h->b1_tensor == 2 && n_groups >= 2));
    if (tensor_path && c.has_empty)
      if (int e = check_cuda(h, cudaMemsetAsync(group_scores, 0, size_t(n_groups_padded) * size_t(ld) * 4, s), "score: memset")) return e;
    cudaEventRecord(h->ev0[slot], s);
    if (tensor_path) {
      if (int e = launch_score_b1_umma(h, q_packed, static_cast<const int32_t*>(h->meta_c.p), n_groups, group_scores, ld, s)) return e;
    } else {
      if (int e = launch_score_b1(h, cand_ids, n_cand, q_packed, n_groups, static_cast<const int32_t*>(h->meta_c.p),
                                  group_scores, ld, s))
        return e;
    }
    cudaEventRecord(h->ev1[slot], s);
    h->ev_count++;
    // ntok is pageable host memory: the async upload staged it before returning, nothing else to wait for
    return B200MS_OK;
  }
  // pages with zero rows (and unused candidate slots) are never touched by the tile kernel: they score 0
  if (c.has_empty || cand_ids) {
    if (int e = check_cuda(h, cudaMemsetAsync(group_scores, 0, size_t(n_groups_padded) * size_t(ld) * 4, s), "score: memset")) return e;
  }
  UnitPlan plan;
  if (cand_ids) {
    if (int e = reserve(h, h->cand_start, size_t(n_cand) * 4)) return e;
    if (int e = reserve(h, h->cand_end, size_t(n_cand) * 4)) return e;
    if (int e = launch_cand_units(h, cand_ids, n_cand, static_cast<int32_t*>(h->cand_start.p),
                                  static_cast<int32_t*>(h->cand_end.p), nullptr, s))
      return e;
    plan.start = static_cast<const int32_t*>(h->cand_start.p);
    plan.end = static_cast<const int32_t*>(h->cand_end.p);
    plan.n_units = n_cand;
    plan.slot_mode = 1;
  }
  cudaEventRecord(h->ev0[slot], s);
  if (int e = launch_score_umma(h, cand_ids ? &plan : nullptr, q_packed, n_groups, group_scores, ld, s)) return e;
  cudaEventRecord(h->ev1[slot], s);
  h->ev_count++;
  return B200MS_OK;
}

B200MS_API int b200ms_score(b200ms_t* h, const void* q_packed, int n_groups, const int32_t* q_lens,
                            const int32_t* group_offsets, int n_q, void* group_scores, int64_t ld, void* stream) {
  if (!h) return B200MS_EINVAL;
  DeviceGuard g(h->device);
  return score_impl(h, q_packed, n_groups, q_lens, group_offsets, n_q, group_scores, ld, nullptr, 0, static_cast<cudaStream_t>(stream));
}

static int score_time_of(b200ms_t* h, int64_t idx, float* ms) {
  const int slot = int(idx % b200ms_t::kEvRing);
  if (cudaEventSynchronize(h->ev1[slot]) != cudaSuccess) return check_cuda(h, cudaGetLastError(), "cudaEventSynchronize");
  if (cudaEventElapsedTime(ms, h->ev0[slot], h->ev1[slot]) != cudaSuccess) return check_cuda(h, cudaGetLastError(), "cudaEventElapsedTime");
  return B200MS_OK;
}

B200MS_API float b200ms_last_score_ms(b200ms_t* h) {
  if (!h) return float(B200MS_EINVAL);
  if (h->ev_count == 0) return float(set_error(h, B200MS_ESTATE, "last_score_ms: no scoring call recorded"));
  DeviceGuard g(h->device);
  float ms = 0.f;
  if (int e = score_time_of(h, h->ev_count - 1, &ms)) return float(e);
  return ms;
}

B200MS_API int64_t b200ms_score_call_count(const b200ms_t* h) { return h ? h->ev_count : 0; }

B200MS_API int b200ms_score_times_ms(b200ms_t* h, float* out_ms, int n) {
  if (!h || !out_ms || n < 0) return B200MS_EINVAL;
  if (n > b200ms_t::kEvRing || n > h->ev_count) return set_error(h, B200MS_EINVAL, "score_times_ms: n exceeds the recorded history (ring of 256)");
  DeviceGuard g(h->device);
  for (int i = 0; i < n; ++i)
    if (int e = score_time_of(h, h->ev_count - n + i, &out_ms[i])) return e;
  return n;
}

B200MS_API int b200ms_topk(b200ms_t* h, const void* group_scores, int score_dtype, int64_t n_pages, int64_t ld,
                           const int32_t* group_offsets, int n_q, const uint32_t* allow_mask, int k, float scale,
                           int64_t id_base, float* top_scores, int64_t* top_ids, int32_t* top_counts, void* stream) {
  if (!h) return B200MS_EINVAL;
  if ((score_dtype != B200MS_F32 && score_dtype != B200MS_I32) || n_pages < 0 || ld < n_pages || n_q < 0 || k < 1 ||
      k > B200MS_MAX_K || (n_q > 0 && (!group_offsets || !top_scores || !top_ids || !top_counts)) ||
      (n_q > 0 && n_pages > 0 && !group_scores))
    return set_error(h, B200MS_EINVAL, "topk: bad arguments (1 <= k <= 4096)");
  if (n_q == 0) return B200MS_OK;
  DeviceGuard g(h->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int e = upload(h, h->meta_a, group_offsets, size_t(n_q + 1) * 4, s)) return e;
  return launch_topk(h, group_scores, score_dtype, n_pages, ld, static_cast<const int32_t*>(h->meta_a.p), n_q, allow_mask,
                     k, scale, id_base, nullptr, top_scores, top_ids, top_counts, s);
}

B200MS_API int b200ms_merge_topk(b200ms_t* h, const float* cand_scores, const int64_t* cand_ids, int n_q, int m, int k,
                                 float* top_scores, int64_t* top_ids, int32_t* top_counts, void* stream) {
  if (!h) return B200MS_EINVAL;
  if (n_q < 0 || m < 1 || m > 2 * B200MS_MAX_K || k < 1 || k > B200MS_MAX_K ||
      (n_q > 0 && (!cand_scores || !cand_ids || !top_scores || !top_ids || !top_counts)))
    return set_error(h, B200MS_EINVAL, "merge_topk: bad arguments (m <= 8192, k <= 4096)");
  DeviceGuard g(h->device);
  return launch_merge_topk(h, cand_scores, cand_ids, n_q, m, k, top_scores, top_ids, top_counts,
                           static_cast<cudaStream_t>(stream));
}

static int search_impl(b200ms_t* h, const void* q_dev, int src_dtype, const int32_t* q_lens, int n_q, int k,
                       const uint32_t* allow_dev, float i8_q_scale, float score_scale, int64_t id_base, float* ts,
                       int64_t* ti, int32_t* tc, cudaStream_t s, const int64_t* cand_ids = nullptr, int n_cand = 0,
                       const int32_t* mask_index_dev = nullptr, int64_t mask_stride = 0) {
  const Corpus& c = h->corpus;
  if (c.dtype < 0) return set_error(h, B200MS_ESTATE, "search: no corpus attached (call b200ms_set_corpus first)");
  if (n_q < 1 || !q_lens || !q_dev || k < 1 || k > B200MS_MAX_K || !ts || !ti || !tc)
    return set_error(h, B200MS_EINVAL, "search: bad arguments (n_q >= 1, 1 <= k <= 4096)");
  const int64_t groups = b200ms_query_groups(q_lens, n_q);
  const int64_t groups_padded = (groups + 3) & ~int64_t(3);
  const int64_t n_items = cand_ids ? n_cand : c.n_pages;
  const int64_t ld = (n_items + 31) & ~int64_t(31);
  if (int e = reserve(h, h->q_packed, size_t(groups_padded > 0 ? groups_padded : 4) * kGroup * size_t(b200ms_row_bytes(c.dtype)))) return e;
  if (int e = reserve(h, h->scores, size_t(groups_padded > 0 ? groups_padded : 4) * size_t(ld > 0 ? ld : 32) * 4)) return e;
  std::vector<int32_t> goff(size_t(n_q) + 1);
  int ng = 0;
  if (int e = b200ms_pack_queries(h, q_dev, src_dtype, q_lens, n_q, h->q_packed.p, c.dtype, i8_q_scale, goff.data(), &ng, s)) return e;
  if (int e = score_impl(h, h->q_packed.p, ng, q_lens, goff.data(), n_q, h->scores.p, ld, cand_ids, n_cand, s)) return e;
  if (ng == 0 || n_items == 0) {
    // nothing scored: every page (if any) has score 0 -- still a defined ranking
    if (int e = check_cuda(h, cudaMemsetAsync(h->scores.p, 0, size_t(groups_padded > 0 ? groups_padded : 4) * size_t(ld > 0 ? ld : 32) * 4, s), "search: memset")) return e;
  }
  const int sdt = c.dtype == B200MS_BF16 ? B200MS_F32 : B200MS_I32;
  if (!cand_ids) {
    if (int e = upload(h, h->meta_a, goff.data(), size_t(n_q + 1) * 4, s)) return e;
    return launch_topk(h, h->scores.p, sdt, c.n_pages, ld, static_cast<const int32_t*>(h->meta_a.p), n_q, allow_dev, k,
                       score_scale, id_base, nullptr, ts, ti, tc, s, mask_index_dev, mask_stride);
  }
  // candidate mode: rank the slots (ties -> lower slot = better first-stage rank), report the page ids; unused slots masked
  if (int e = reserve(h, h->cand_mask, size_t((n_cand + 31) / 32) * 4)) return e;
  if (int e = launch_cand_units(h, cand_ids, n_cand, nullptr, nullptr, static_cast<uint32_t*>(h->cand_mask.p), s)) return e;
  if (int e = upload(h, h->meta_a, goff.data(), size_t(n_q + 1) * 4, s)) return e;
  return launch_topk(h, h->scores.p, sdt, n_cand, ld, static_cast<const int32_t*>(h->meta_a.p), n_q,
                     static_cast<const uint32_t*>(h->cand_mask.p), k, score_scale, 0, cand_ids, ts, ti, tc, s);
}

B200MS_API int b200ms_rerank_device(b200ms_t* h, const void* q_dev, int src_dtype, const int32_t* q_lens, int n_q,
                                    const int64_t* cand_ids_dev, int n_cand, int k, float i8_q_scale, float score_scale,
                                    float* top_scores_dev, int64_t* top_ids_dev, int32_t* top_counts_dev, void* stream) {
  if (!h) return B200MS_EINVAL;
  if (!src_dtype_ok(src_dtype) || !cand_ids_dev || n_cand < 1)
    return set_error(h, B200MS_EINVAL, "rerank_device: bad arguments (F32/BF16 queries, n_cand >= 1)");
  DeviceGuard g(h->device);
  return search_impl(h, q_dev, src_dtype, q_lens, n_q, k, nullptr, i8_q_scale, score_scale, 0, top_scores_dev, top_ids_dev,
                     top_counts_dev, static_cast<cudaStream_t>(stream), cand_ids_dev, n_cand);
}

B200MS_API int b200ms_search_device(b200ms_t* h, const void* q_dev, int src_dtype, const int32_t* q_lens, int n_q, int k,
                                    const uint32_t* allow_mask_dev, float i8_q_scale, float score_scale, int64_t id_base,
                                    float* top_scores_dev, int64_t* top_ids_dev, int32_t* top_counts_dev, void* stream) {
  if (!h) return B200MS_EINVAL;
  if (!src_dtype_ok(src_dtype)) return set_error(h, B200MS_EINVAL, "search_device: src dtype must be F32 or BF16");
  DeviceGuard g(h->device);
  return search_impl(h, q_dev, src_dtype, q_lens, n_q, k, allow_mask_dev, i8_q_scale, score_scale, id_base, top_scores_dev,
                     top_ids_dev, top_counts_dev, static_cast<cudaStream_t>(stream));
}

B200MS_API int b200ms_search_device_masked(b200ms_t* h, const void* q_dev, int src_dtype, const int32_t* q_lens, int n_q,
                                           int k, const uint32_t* allow_masks_dev, int n_masks,
                                           const int32_t* mask_index_dev, float i8_q_scale, float score_scale,
                                           int64_t id_base, float* top_scores_dev, int64_t* top_ids_dev,
                                           int32_t* top_counts_dev, void* stream) {
  if (!h) return B200MS_EINVAL;
  if (!src_dtype_ok(src_dtype)) return set_error(h, B200MS_EINVAL, "search_device_masked: src dtype must be F32 or BF16");
  if (n_masks < 0 || (n_masks > 0 && (!allow_masks_dev || !mask_index_dev)))
    return set_error(h, B200MS_EINVAL, "search_device_masked: n_masks > 0 needs the mask matrix and one index per query");
  DeviceGuard g(h->device);
  return search_impl(h, q_dev, src_dtype, q_lens, n_q, k, n_masks > 0 ? allow_masks_dev : nullptr, i8_q_scale, score_scale,
                     id_base, top_scores_dev, top_ids_dev, top_counts_dev, static_cast<cudaStream_t>(stream), nullptr, 0,
                     n_masks > 0 ? mask_index_dev : nullptr, (h->corpus.n_pages + 31) / 32);
}

static int search_host_impl(b200ms_t* h, const float* q_host, const int32_t* q_lens, int n_q, int k,
                            const uint32_t* masks_host, int n_masks, const int32_t* mask_index_host, float i8_q_scale,
                            float score_scale, int64_t id_base, float* top_scores_host, int64_t* top_ids_host,
                            int32_t* top_counts_host) {
  if (n_q < 1 || !q_lens || !q_host || k < 1 || k > B200MS_MAX_K || !top_scores_host || !top_ids_host || !top_counts_host)
    return set_error(h, B200MS_EINVAL, "search_host: bad arguments (n_q >= 1, 1 <= k <= 4096)");
  DeviceGuard g(h->device);
  cudaStream_t s = h->stream;
  int64_t rows = 0;
  for (int i = 0; i < n_q; ++i) rows += q_lens[i] > 0 ? q_lens[i] : 0;
  if (int e = reserve(h, h->q_raw, size_t(rows > 0 ? rows : 1) * kDim * 4)) return e;
  if (int e = reserve(h, h->out_s, size_t(n_q) * k * 4)) return e;
  if (int e = reserve(h, h->out_i, size_t(n_q) * k * 8)) return e;
  if (int e = reserve(h, h->out_c, size_t(n_q) * 4)) return e;
  if (rows > 0)
    if (int e = check_cuda(h, cudaMemcpyAsync(h->q_raw.p, q_host, size_t(rows) * kDim * 4, cudaMemcpyHostToDevice, s), "search_host: H2D queries")) return e;
  const uint32_t* allow_dev = nullptr;
  const int32_t* index_dev = nullptr;
  const int64_t words = (h->corpus.n_pages + 31) / 32;
  if (masks_host && n_masks > 0 && h->corpus.n_pages > 0) {
    if (int e = upload(h, h->mask, masks_host, size_t(n_masks) * size_t(words) * 4, s)) return e;
    allow_dev = static_cast<const uint32_t*>(h->mask.p);
    if (mask_index_host) {
      for (int i = 0; i < n_q; ++i)
        if (mask_index_host[i] < -1 || mask_index_host[i] >= n_masks)
          return set_error(h, B200MS_EINVAL, "search_host_masked: mask_index out of range");
      if (int e = upload(h, h->mask_index, mask_index_host, size_t(n_q) * 4, s)) return e;
      index_dev = static_cast<const int32_t*>(h->mask_index.p);
    }
  }
  if (int e = search_impl(h, h->q_raw.p, B200MS_F32, q_lens, n_q, k, allow_dev, i8_q_scale, score_scale, id_base,
                          static_cast<float*>(h->out_s.p), static_cast<int64_t*>(h->out_i.p),
                          static_cast<int32_t*>(h->out_c.p), s, nullptr, 0, index_dev, words))
    return e;
  cudaMemcpyAsync(top_scores_host, h->out_s.p, size_t(n_q) * k * 4, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(top_ids_host, h->out_i.p, size_t(n_q) * k * 8, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(top_counts_host, h->out_c.p, size_t(n_q) * 4, cudaMemcpyDeviceToHost, s);
  return check_cuda(h, cudaStreamSynchronize(s), "search_host: stream sync");
}

B200MS_API int b200ms_search_host(b200ms_t* h, const float* q_host, const int32_t* q_lens, int n_q, int k,
                                  const uint32_t* allow_mask_host, float i8_q_scale, float score_scale, int64_t id_base,
                                  float* top_scores_host, int64_t* top_ids_host, int32_t* top_counts_host) {
  if (!h) return B200MS_EINVAL;
  return search_host_impl(h, q_host, q_lens, n_q, k, allow_mask_host, allow_mask_host ? 1 : 0, nullptr, i8_q_scale,
                          score_scale, id_base, top_scores_host, top_ids_host, top_counts_host);
}

B200MS_API int b200ms_search_host_masked(b200ms_t* h, const float* q_host, const int32_t* q_lens, int n_q, int k,
                                         const uint32_t* allow_masks_host, int n_masks, const int32_t* mask_index_host,
                                         float i8_q_scale, float score_scale, int64_t id_base, float* top_scores_host,
                                         int64_t* top_ids_host, int32_t* top_counts_host) {
  if (!h) return B200MS_EINVAL;
  if (n_masks < 0 || (n_masks > 0 && (!allow_masks_host || !mask_index_host)))
    return set_error(h, B200MS_EINVAL, "search_host_masked: n_masks > 0 needs the mask matrix and one index per query");
  return search_host_impl(h, q_host, q_lens, n_q, k, allow_masks_host, n_masks, mask_index_host, i8_q_scale, score_scale,
                          id_base, top_scores_host, top_ids_host, top_counts_host);
}

// ------------------------------------------------------------------------------------------------ FDE (next row f-1)
B200MS_API int b200ms_fde_configure(b200ms_t* h, int reps, int ksim, int proj_dim, float scale, const float* simhash,
                                    const int32_t* ams_index, const float* ams_sign) {
  if (!h) return B200MS_EINVAL;
  if (reps < 1 || ksim < 1 || ksim > 8 || proj_dim < 1 || proj_dim > 64 || (proj_dim << ksim) > 512 || ((proj_dim << ksim) % 8) ||
      !simhash || !ams_index || !ams_sign)
    return set_error(h, B200MS_EINVAL, "fde_configure: need 1<=ksim<=8, proj_dim<=64, proj_dim*2^ksim <= 512 and a multiple of 8");
  for (int i = 0; i < reps * kDim; ++i)
    if (ams_index[i] < 0 || ams_index[i] >= proj_dim) return set_error(h, B200MS_EINVAL, "fde_configure: ams_index out of range");
  DeviceGuard g(h->device);
  cudaStream_t s = h->stream;
  if (int e = upload(h, h->fde_simhash, simhash, size_t(reps) * kDim * ksim * 4, s)) return e;
  if (int e = upload(h, h->fde_ams_index, ams_index, size_t(reps) * kDim * 4, s)) return e;
  if (int e = upload(h, h->fde_ams_sign, ams_sign, size_t(reps) * kDim * 4, s)) return e;
  if (int e = check_cuda(h, cudaStreamSynchronize(s), "fde_configure: sync")) return e;
  h->fde_reps = reps;
  h->fde_ksim = ksim;
  h->fde_proj = proj_dim;
  h->fde_scale = scale;
  h->fde_dim = reps * (1 << ksim) * proj_dim;
  return B200MS_OK;
}

B200MS_API int64_t b200ms_fde_dim(const b200ms_t* h) { return h ? h->fde_dim : 0; }

B200MS_API int b200ms_fde_encode(b200ms_t* h, const void* rows, int src_dtype, const int32_t* item_lens, int64_t n_items,
                                 int is_document, float* out, void* stream) {
  if (!h) return B200MS_EINVAL;
  if (h->fde_dim == 0) return set_error(h, B200MS_ESTATE, "fde_encode: call b200ms_fde_configure first");
  if (!src_dtype_ok(src_dtype) || n_items < 0 || n_items > 65535 || (n_items > 0 && (!rows || !item_lens || !out)))
    return set_error(h, B200MS_EINVAL, "fde_encode: bad arguments (at most 65535 items per call)");
  if (n_items == 0) return B200MS_OK;
  DeviceGuard g(h->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::vector<int64_t> st(size_t(n_items) + 1);
  st[0] = 0;
  for (int64_t i = 0; i < n_items; ++i) st[i + 1] = st[i] + (item_lens[i] > 0 ? item_lens[i] : 0);
  if (int e = upload(h, h->meta_a, st.data(), st.size() * 8, s)) return e;
  return launch_fde_encode(h, rows, src_dtype, static_cast<const int64_t*>(h->meta_a.p), int(n_items), is_document, out, s);
}

B200MS_API int b200ms_fde_finalize(b200ms_t* h, const float* fde, int64_t n, void* out_rows, float* inv_norm, void* stream) {
  if (!h) return B200MS_EINVAL;
  if (h->fde_dim == 0) return set_error(h, B200MS_ESTATE, "fde_finalize: call b200ms_fde_configure first");
  if (n < 0 || (n > 0 && (!fde || !out_rows || !inv_norm))) return set_error(h, B200MS_EINVAL, "fde_finalize: bad arguments");
  DeviceGuard g(h->device);
  return launch_fde_finalize(h, fde, n, out_rows, inv_norm, static_cast<cudaStream_t>(stream));
}

B200MS_API int b200ms_fde_scan(b200ms_t* h, const void* fde_rows, const float* inv_norm, int64_t n_pages, const float* q_fde,
                               int n_q, float* scores, int64_t ld, void* stream) {
  if (!h) return B200MS_EINVAL;
  if (h->fde_dim == 0) return set_error(h, B200MS_ESTATE, "fde_scan: call b200ms_fde_configure first");
  if (n_pages < 0 || n_q < 0 || ld < n_pages || (n_pages > 0 && n_q > 0 && (!fde_rows || !inv_norm || !q_fde || !scores)) ||
      (reinterpret_cast<uintptr_t>(fde_rows) & 15))
    return set_error(h, B200MS_EINVAL, "fde_scan: bad arguments (rows 16-byte aligned, ld >= n_pages)");
  DeviceGuard g(h->device);
  return launch_fde_scan(h, fde_rows, inv_norm, n_pages, q_fde, n_q, scores, ld, static_cast<cudaStream_t>(stream));
}
