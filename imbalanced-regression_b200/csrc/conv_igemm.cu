// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM + TMA),
// NHWC bf16 activations, fp32 accumulation.  Stage (i) of the hot path:
// replaces the cuDNN convolutions behind agedb-dir/resnet.py:46-51,79,112-118
// (forward) and their autograd backward (data and weight gradients).
//
// One CTA computes one 128 x BN output tile:
//   warps 0-3  gather the A operand (im2col rows) global -> swizzled smem with
//              16-byte cp.async (zero-fill = padding), then run the epilogue
//              (tcgen05.ld TMEM -> registers -> global);
//   warp 4     streams the B operand (weights, or dY for wgrad) with TMA;
//   warp 5     owns TMEM and issues tcgen05.mma from one elected thread.
// A ring of STAGES smem slots is handed between them with mbarriers.
//
//   FPROP  Y[p, co]  = sum_k  A[p, k] W[co, k]      A, W K-major      (k = (r, s, c))
//   DGRAD  dX[p, c]  = sum_k  A'[p, k] Wt[c, k]     same kernel, transposed gather (k = (r, s, co))
//   WGRAD  dW[k, co] = sum_p  A[p, k] dY[p, co]     both operands MN-major, K = pixels, split-K partials
#include "common.cuh"
#include "tc.cuh"
#include "conv.cuh"

namespace dirb200 {
using namespace tc;

constexpr int BM = 128;  // GEMM rows per tile (pixels; for wgrad: 2 chunks x 64 gathered channels)
constexpr int BK = 64;   // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kProducerThreads = 128;
constexpr int kThreads = 192;

struct IgemmParams {
  const __nv_bfloat16* src;  // gathered tensor, NHWC
  int n, hs, ws, cs;         // its shape
  int hm, wm;                // pixel grid that indexes GEMM rows (fprop/wgrad: conv output grid; dgrad: conv input grid)
  int kh, kw, stride, pad;
  int transposed;            // 0: fprop-style gather, 1: dgrad-style (source = dY of a strided conv)
  int cpb;                   // 64-channel blocks per filter tap (cs / 64); stem: unused
  long long pixels;          // n * hm * wm
  int num_kblocks;           // fprop/dgrad: kh*kw*cpb ; wgrad: ceil(pixels / 64)
  int kblocks_per_split;     // wgrad split-K
  int total_chunks;          // wgrad: kh*kw*cpb (64-row chunks of the K_total x Cout result)
  int n_tiles;               // tiles along GEMM N
  int ldc;                   // output row stride in elements (Cout for fprop, Cin for dgrad, Cout for wgrad partials)
  void* out;                 // bf16 [pixels][ldc]   or   fp32 [splits][total_chunks*64][ldc]
};

template <int BN>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 128) ? 3 : 4;   // ~97 KB per CTA -> two CTAs per SM overlap each other's prologue/epilogue
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kSmemBytes = kBarOffset + (2 * kStages + 2) * 8 + 1024;
  static constexpr int kTmemCols = BN;  // power of two >= 32
};

// ---- A-operand gather --------------------------------------------------------------------------------
// A tile row = 128 contiguous bytes of the source tensor (64 channels of one pixel; stem: 4 taps x 16 channels).
// The 8 lanes of a quarter-warp copy the 8 x 16 B of one row, so every warp-wide cp.async touches 4 full 128-byte
// lines (fully coalesced L2 requests); a thread therefore serves 8 different rows, always the same 16-byte column.
//
// packed pixel: bit 31 valid | n (13 bits) << 18 | y (9 bits) << 9 | x (9 bits)
__device__ __forceinline__ uint32_t pack_pixel(long long p, const IgemmParams& P) {
  if (p >= P.pixels) return 0u;
  const uint32_t pp = static_cast<uint32_t>(p);
  const uint32_t hw = static_cast<uint32_t>(P.hm * P.wm);
  const uint32_t n = pp / hw;
  const uint32_t rem = pp - n * hw;
  const uint32_t y = rem / static_cast<uint32_t>(P.wm);
  const uint32_t x = rem - y * static_cast<uint32_t>(P.wm);
  return 0x80000000u | (n << 18) | (y << 9) | x;
}

// Row-invariant part of the gather address, unpacked once: image base row n*hs, and the tap-0 coordinates
// (fprop-style: y*stride - pad ; dgrad-style: y + pad).
struct RowPre {
  int nb, yb, xb;   // yb == INT_MIN/2 marks an invalid (out-of-range) row
};
__device__ __forceinline__ RowPre row_pre(const IgemmParams& P, uint32_t pk) {
  RowPre rp;
  const int n = (pk >> 18) & 0x1FFF, y = (pk >> 9) & 0x1FF, x = pk & 0x1FF;
  rp.nb = n * P.hs;
  if (!P.transposed) {
    rp.yb = y * P.stride - P.pad;
    rp.xb = x * P.stride - P.pad;
  } else {
    rp.yb = y + P.pad;
    rp.xb = x + P.pad;
  }
  if (!(pk >> 31)) rp.yb = -(1 << 28);
  return rp;
}

// Source of the 16 bytes at channel offset `coff` of filter tap (r, s).
// fprop-style: input pixel (y*stride - pad + r, x*stride - pad + s).
// dgrad-style: the conv-output pixel (ho, wo) with ho*stride - pad + r == y (stride 1 or 2; must divide exactly).
__device__ __forceinline__ const __nv_bfloat16* tap_source(const IgemmParams& P, const RowPre& rp, int r, int s,
                                                           int coff, bool& ok) {
  int hi, wi;
  ok = true;
  if (!P.transposed) {
    hi = rp.yb + r;
    wi = rp.xb + s;
  } else {
    hi = rp.yb - r;
    wi = rp.xb - s;
    if (P.stride == 2) {
      ok = ((hi | wi) & 1) == 0;
      hi >>= 1;
      wi >>= 1;
    }
  }
  ok = ok && static_cast<unsigned>(hi) < static_cast<unsigned>(P.hs) &&
       static_cast<unsigned>(wi) < static_cast<unsigned>(P.ws);
  return ok ? P.src + (static_cast<size_t>(rp.nb + hi) * P.ws + wi) * P.cs + coff : P.src;
}

template <int BN, bool WGRAD, bool STEM>
__global__ void __launch_bounds__(kThreads, 2)
igemm_kernel(const __grid_constant__ CUtensorMap tmap_b, const IgemmParams P) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::kBarOffset;
  auto a_addr = [&](int s) { return smem_base + s * C::kStageBytes; };
  auto b_addr = [&](int s) { return smem_base + s * C::kStageBytes + C::kABytes; };
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * C::kStages);
  const uint32_t tmem_holder = bar_base + 8u * (2 * C::kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % P.n_tiles;
  const int m_tile = blockIdx.x / P.n_tiles;
  const int split = blockIdx.y;
  int kb_begin = 0, kb_end = P.num_kblocks;
  if constexpr (WGRAD) {
    kb_begin = split * P.kblocks_per_split;
    kb_end = min(P.num_kblocks, kb_begin + P.kblocks_per_split);
  }
  const int nk = max(0, kb_end - kb_begin);

  if (warp == 5) {
    if (lane == 0) {
      for (int s = 0; s < C::kStages; ++s) {
        mbar_init(full_bar(s), kProducerThreads + 1);
        mbar_init(empty_bar(s), 1);
      }
      mbar_init(accum_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_holder, C::kTmemCols);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_holder));

  if (warp < 4) {
    // ============================ A producer ============================
    const int j = lane & 7;          // 16-byte column served by this thread
    const int q = lane >> 3;         // row within a group of 4
    RowPre rows8[8];                 // fprop/dgrad: the 8 rows this thread serves
    int chunk_r = 0, chunk_s = 0, chunk_c0 = 0;
    bool chunk_ok = true;
    uint32_t tile_off = 0;
    if constexpr (!WGRAD) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        rows8[i] = row_pre(P, pack_pixel(static_cast<long long>(m_tile) * BM + warp * 32 + 4 * i + q, P));
      tile_off = warp * 32 * 128;
    } else {
      // warp = (chunk, half): 64 gathered channels x 32 of the 64 pixel rows of the k-block
      const int chunk = warp >> 1;
      const int gchunk = m_tile * 2 + chunk;           // 64-row chunk of the [K_total, Cout] result
      chunk_ok = gchunk < P.total_chunks;
      if constexpr (STEM) {
        chunk_r = gchunk;                              // filter row r'
      } else {
        const int tap = gchunk / P.cpb;
        chunk_c0 = (gchunk - tap * P.cpb) * 64;
        chunk_r = tap / P.kw;
        chunk_s = tap - chunk_r * P.kw;
      }
      tile_off = chunk * 8192 + (warp & 1) * 32 * 128;
    }
    for (int it = 0; it < nk; ++it) {
      const int s = it % C::kStages;
      const uint32_t ph = (it / C::kStages) & 1;
      mbar_wait(empty_bar(s), ph ^ 1u);
      const int kb = kb_begin + it;
      int r, sx, coff;
      uint32_t mypk = 0;
      if constexpr (!WGRAD) {
        if constexpr (STEM) {
          r = kb; sx = j >> 1; coff = (j & 1) * 8;
        } else {
          const int tap = kb / P.cpb;
          coff = (kb - tap * P.cpb) * 64 + j * 8;
          r = tap / P.kw;
          sx = tap - r * P.kw;
        }
      } else {
        r = chunk_r;
        if constexpr (STEM) { sx = j >> 1; coff = (j & 1) * 8; } else { sx = chunk_s; coff = chunk_c0 + j * 8; }
        mypk = chunk_ok ? pack_pixel(static_cast<long long>(kb) * 64 + (warp & 1) * 32 + lane, P) : 0u;
      }
      const uint32_t dst_base = a_addr(s) + tile_off;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = 4 * i + q;                     // row within this warp's 32 rows
        RowPre rp;
        if constexpr (!WGRAD) rp = rows8[i];
        else rp = row_pre(P, __shfl_sync(0xffffffffu, mypk, row));
        bool ok;
        const __nv_bfloat16* src = tap_source(P, rp, r, sx, coff, ok);
        cp_async16(dst_base + row * 128 + ((j ^ (row & 7)) << 4), src, ok ? 16u : 0u);
      }
      // the mbarrier receives this thread's arrival when all of its cp.async above have landed (no wait here:
      // the ring depth alone bounds the loads in flight), as CUTLASS's sm100 cp.async->UMMA mainloop does
      cp_async_mbar_arrive_noinc(full_bar(s));
    }

    // ============================== epilogue ==============================
    mbar_wait(accum_bar, 0);
    tcgen05_fence_after();
    const int row = warp * 32 + lane;
    const int n0 = n_tile * BN;
    if constexpr (!WGRAD) {
      const long long p = static_cast<long long>(m_tile) * BM + row;
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(P.out) + p * P.ldc + n0;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        if (p < P.pixels && nk > 0) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            pk[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          uint4* o = reinterpret_cast<uint4*>(out + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
      }
    } else {
      const int krow = m_tile * BM + row;            // row of the [K_total, Cout] result
      const int ktot = P.total_chunks * 64;
      float* out = reinterpret_cast<float*>(P.out) + (static_cast<size_t>(split) * ktot + krow) * P.ldc + n0;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        if (krow < ktot) {
          float4* o = reinterpret_cast<float4*>(out + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 t;
            if (nk > 0)
              t = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                              __uint_as_float(v[4 * j + 3]));
            else
              t = make_float4(0.f, 0.f, 0.f, 0.f);
            o[j] = t;
          }
        }
      }
    }
    tcgen05_fence_before();
  } else if (warp == 4) {
    // ============================ B producer (TMA) ============================
    if (lane == 0) {
      const int n0 = n_tile * BN;
      for (int it = 0; it < nk; ++it) {
        const int s = it % C::kStages;
        const uint32_t ph = (it / C::kStages) & 1;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), C::kBBytes);
        const int kb = kb_begin + it;
        if constexpr (!WGRAD) {
          tma_load_2d(b_addr(s), &tmap_b, full_bar(s), kb * BK, n0);
        } else {
#pragma unroll
          for (int i = 0; i < BN / 64; ++i) tma_load_2d(b_addr(s) + i * 8192, &tmap_b, full_bar(s), n0 + 64 * i, kb * 64);
        }
      }
    }
  } else {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN, WGRAD ? 1 : 0, WGRAD ? 1 : 0);
      for (int it = 0; it < nk; ++it) {
        const int s = it % C::kStages;
        const uint32_t ph = (it / C::kStages) & 1;
        mbar_wait(full_bar(s), ph);
        tcgen05_fence_after();
        // K-major: 8-row atoms 1024 B apart; MN-major: 64-wide chunks 8192 B apart (LBO), 8-k atoms 1024 B (SBO)
        const uint64_t adesc = make_smem_desc(a_addr(s), WGRAD ? 8192u : 16u, 1024u);
        const uint64_t bdesc = make_smem_desc(b_addr(s), WGRAD ? 8192u : 16u, 1024u);
        constexpr uint32_t kadv = WGRAD ? (2048u >> 4) : (32u >> 4);   // one UMMA_K (=16) step, in 16-byte units
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tmem_base, adesc + static_cast<uint64_t>(k * kadv), bdesc + static_cast<uint64_t>(k * kadv), idesc,
                    (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(empty_bar(s));
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == 5) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// --------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor [outer][inner] (inner contiguous), box = 64 x box_outer, 128-byte swizzle.
int make_tmap_bf16_2d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return DIRB200_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu stride=%llu", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes);
    return DIRB200_ERR_CUDA;
  }
  return DIRB200_OK;
}

template <int BN, bool WGRAD, bool STEM>
static int launch_igemm(const CUtensorMap& tm, const IgemmParams& P, int m_tiles, int splits, cudaStream_t st) {
  using C = Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    DIRB_CUDA(cudaFuncSetAttribute(igemm_kernel<BN, WGRAD, STEM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   C::kSmemBytes));
    configured = true;
  }
  dim3 grid(static_cast<unsigned>(m_tiles * P.n_tiles), static_cast<unsigned>(splits));
  igemm_kernel<BN, WGRAD, STEM><<<grid, kThreads, C::kSmemBytes, st>>>(tm, P);
  DIRB_LAUNCHED();
  return DIRB200_OK;
}

static int check_shape(const ConvShape& s, bool stem, const char* who) {
  DIRB_CHECK_ARG(s.n > 0 && s.h > 0 && s.w > 0 && s.kh > 0 && s.kw > 0 && s.stride > 0 && s.pad >= 0,
                 "%s: bad conv shape", who);
  DIRB_CHECK_ARG(s.cout % 64 == 0, "%s: Cout must be a multiple of 64 (got %d)", who, s.cout);
  if (stem)
    DIRB_CHECK_ARG(s.cin == 16 && s.kh == 4 && s.kw == 4 && s.stride == 1, "%s: stem expects the 4x4x16 s2d form", who);
  else
    DIRB_CHECK_ARG(s.cin % 64 == 0, "%s: Cin must be a multiple of 64 (got %d)", who, s.cin);
  return DIRB200_OK;
}

// Y[n,ho,wo,cout] = conv(X[n,h,w,cin], W[cout][kh][kw][cin])
int conv_fprop(const __nv_bfloat16* x, const __nv_bfloat16* w, __nv_bfloat16* y, const ConvShape& s, bool stem,
               cudaStream_t st) {
  if (int rc = check_shape(s, stem, "conv_fprop")) return rc;
  const int ktot = s.kh * s.kw * s.cin;
  IgemmParams P{};
  P.src = x; P.n = s.n; P.hs = s.h; P.ws = s.w; P.cs = s.cin; P.hm = s.ho; P.wm = s.wo;
  P.kh = s.kh; P.kw = s.kw; P.stride = s.stride; P.pad = s.pad; P.transposed = 0;
  P.cpb = stem ? 1 : s.cin / 64;
  P.pixels = static_cast<long long>(s.n) * s.ho * s.wo;
  P.num_kblocks = ktot / 64;
  P.ldc = s.cout; P.out = y;
  const int bn = (s.cout % 128 == 0) ? 128 : 64;
  P.n_tiles = s.cout / bn;
  const int m_tiles = static_cast<int>((P.pixels + BM - 1) / BM);
  CUtensorMap tm;
  if (int rc = make_tmap_bf16_2d(&tm, w, ktot, s.cout, static_cast<uint64_t>(ktot) * 2, bn)) return rc;
  if (stem) return bn == 128 ? launch_igemm<128, false, true>(tm, P, m_tiles, 1, st) : launch_igemm<64, false, true>(tm, P, m_tiles, 1, st);
  return bn == 128 ? launch_igemm<128, false, false>(tm, P, m_tiles, 1, st) : launch_igemm<64, false, false>(tm, P, m_tiles, 1, st);
}

// dX[n,h,w,cin] = conv_transpose(dY[n,ho,wo,cout], Wt[cin][kh][kw][cout])
int conv_dgrad(const __nv_bfloat16* dy, const __nv_bfloat16* wt, __nv_bfloat16* dx, const ConvShape& s,
               cudaStream_t st) {
  if (int rc = check_shape(s, false, "conv_dgrad")) return rc;
  DIRB_CHECK_ARG(s.stride == 1 || s.stride == 2, "conv_dgrad: stride must be 1 or 2 (got %d)", s.stride);
  const int ktot = s.kh * s.kw * s.cout;
  IgemmParams P{};
  P.src = dy; P.n = s.n; P.hs = s.ho; P.ws = s.wo; P.cs = s.cout; P.hm = s.h; P.wm = s.w;
  P.kh = s.kh; P.kw = s.kw; P.stride = s.stride; P.pad = s.pad; P.transposed = 1;
  P.cpb = s.cout / 64;
  P.pixels = static_cast<long long>(s.n) * s.h * s.w;
  P.num_kblocks = ktot / 64;
  P.ldc = s.cin; P.out = dx;
  const int bn = (s.cin % 128 == 0) ? 128 : 64;
  P.n_tiles = s.cin / bn;
  const int m_tiles = static_cast<int>((P.pixels + BM - 1) / BM);
  CUtensorMap tm;
  if (int rc = make_tmap_bf16_2d(&tm, wt, ktot, s.cin, static_cast<uint64_t>(ktot) * 2, bn)) return rc;
  return bn == 128 ? launch_igemm<128, false, false>(tm, P, m_tiles, 1, st) : launch_igemm<64, false, false>(tm, P, m_tiles, 1, st);
}

int conv_wgrad_splits(const ConvShape& s) {
  const long long pixels = static_cast<long long>(s.n) * s.ho * s.wo;
  const int kblocks = static_cast<int>((pixels + 63) / 64);
  const int chunks = s.kh * s.kw * s.cin / 64;
  const int bn = (s.cout % 128 == 0) ? 128 : 64;
  const int tiles = ((chunks + 1) / 2) * (s.cout / bn);
  int splits = (2 * num_sms() + tiles - 1) / tiles;
  if (splits < 1) splits = 1;
  const int max_splits = (kblocks + 7) / 8;             // at least 8 k-blocks per split
  if (splits > max_splits) splits = max_splits < 1 ? 1 : max_splits;
  return splits;
}

size_t conv_wgrad_workspace_bytes(const ConvShape& s) {
  return static_cast<size_t>(conv_wgrad_splits(s)) * s.kh * s.kw * s.cin * s.cout * sizeof(float);
}

// partial[split][(r,s,c)][cout] = sum over the split's pixels of X_gathered^T dY
int conv_wgrad_partials(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* partial, const ConvShape& s, bool stem,
                        int* splits_out, cudaStream_t st) {
  if (int rc = check_shape(s, stem, "conv_wgrad")) return rc;
  IgemmParams P{};
  P.src = x; P.n = s.n; P.hs = s.h; P.ws = s.w; P.cs = s.cin; P.hm = s.ho; P.wm = s.wo;
  P.kh = s.kh; P.kw = s.kw; P.stride = s.stride; P.pad = s.pad; P.transposed = 0;
  P.cpb = stem ? 1 : s.cin / 64;
  P.pixels = static_cast<long long>(s.n) * s.ho * s.wo;
  P.num_kblocks = static_cast<int>((P.pixels + 63) / 64);
  P.total_chunks = s.kh * s.kw * s.cin / 64;
  const int splits = conv_wgrad_splits(s);
  P.kblocks_per_split = (P.num_kblocks + splits - 1) / splits;
  P.ldc = s.cout; P.out = partial;
  const int bn = (s.cout % 128 == 0) ? 128 : 64;
  P.n_tiles = s.cout / bn;
  const int m_tiles = (P.total_chunks + 1) / 2;
  *splits_out = splits;
  CUtensorMap tm;
  if (int rc = make_tmap_bf16_2d(&tm, dy, s.cout, static_cast<uint64_t>(P.pixels), static_cast<uint64_t>(s.cout) * 2, 64))
    return rc;
  if (stem) return bn == 128 ? launch_igemm<128, true, true>(tm, P, m_tiles, splits, st) : launch_igemm<64, true, true>(tm, P, m_tiles, splits, st);
  return bn == 128 ? launch_igemm<128, true, false>(tm, P, m_tiles, splits, st) : launch_igemm<64, true, false>(tm, P, m_tiles, splits, st);
}

}  // namespace dirb200
