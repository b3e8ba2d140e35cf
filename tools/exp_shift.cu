// Micro-experiment (not part of the library): can a tcgen05 K-major SWIZZLE_128B A-operand descriptor start at an
// arbitrary 128-byte ROW of a TMA-written tile (start address not 1024-byte aligned)?  That is what an input patch kept
// resident in shared memory across the 9 taps of a 3x3 convolution needs: tap (r, s) reads the same patch displaced by
// (r * padded_width + s) rows.  For every shift s in [0, 24) the kernel runs D = A[s : s + 128, :] * B^T with
//   variant 0: descriptor base_offset = 0,   variant 1: base_offset = (start_address >> 7) & 7
// and the host compares with a CPU reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/exp_shift tools/exp_shift.cu -lcuda && gpurun_out/exp_shift
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../imbalanced-regression_b200/csrc/tc.cuh"

using namespace dirb200::tc;

constexpr int kRowsA = 160, kN = 64, kK = 64, kShifts = 24;

__global__ void __launch_bounds__(128, 1)
shift_kernel(const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tmb, float* out, int variant) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + kRowsA * 128, bar = b_smem + kN * 128, mbar2 = bar + 8, holder = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mbar2, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(holder, 64);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(holder));
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, kRowsA * 128 + kN * 128);
    tma_load_2d(a_smem, &tma, bar, 0, 0);
    tma_load_2d(b_smem, &tmb, bar, 0, 0);
  }
  mbar_wait(bar, 0);
  tcgen05_fence_after();
  constexpr uint32_t idesc = make_idesc(128, kN, 0, 0);
  for (int s = 0; s < kShifts; ++s) {
    if (threadIdx.x == 0) {
      const uint32_t start = a_smem + s * 128;
      uint64_t ad = make_smem_desc(start, 16u, 1024u);
      if (variant == 1) ad |= static_cast<uint64_t>((start >> 7) & 7u) << 49;
      const uint64_t bd = make_smem_desc(b_smem, 16u, 1024u);
      for (int k = 0; k < kK / 16; ++k) umma_bf16(tmem, ad + (k * 2), bd + (k * 2), idesc, k > 0 ? 1u : 0u);
      umma_commit(mbar2);
    }
    mbar_wait(mbar2, s & 1);
    tcgen05_fence_after();
    for (int c = 0; c < kN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + c * 32 + (static_cast<uint32_t>(warp * 32) << 16), v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j)
        out[(static_cast<size_t>(s) * 128 + warp * 32 + lane) * kN + c * 32 + j] = __uint_as_float(v[j]);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static CUtensorMap make_map(void* ptr, int rows, int box_rows) {
  CUtensorMap tm;
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return tm;
}

int main() {
  cudaFree(0);
  std::vector<__nv_bfloat16> ha(kRowsA * kK), hb(kN * kK);
  std::vector<float> fa(kRowsA * kK), fb(kN * kK);
  srand(1);
  for (size_t i = 0; i < ha.size(); ++i) { ha[i] = __float2bfloat16((rand() % 17 - 8) / 8.f); fa[i] = __bfloat162float(ha[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { hb[i] = __float2bfloat16((rand() % 13 - 6) / 4.f); fb[i] = __bfloat162float(hb[i]); }
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, sizeof(float) * kShifts * 128 * kN);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const CUtensorMap tma = make_map(da, kRowsA, kRowsA), tmb = make_map(db, kN, kN);
  const int smem = kRowsA * 128 + kN * 128 + 64 + 1024;
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> ho(kShifts * 128 * kN);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dout, 0, ho.size() * 4);
    shift_kernel<<<1, 128, smem>>>(tma, tmb, dout, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: kernel error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
    printf("variant %d (base_offset %s):", variant, variant ? "= (start>>7)&7" : "= 0");
    for (int s = 0; s < kShifts; ++s) {
      double worst = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < kN; ++n) {
          double ref = 0;
          for (int k = 0; k < kK; ++k) ref += (double)fa[(m + s) * kK + k] * fb[n * kK + k];
          worst = fmax(worst, fabs(ref - ho[(static_cast<size_t>(s) * 128 + m) * kN + n]));
        }
      printf(" s%d:%s", s, worst < 1e-3 ? "ok" : "BAD");
    }
    printf("\n");
  }
  return 0;
}
