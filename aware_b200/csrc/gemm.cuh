// Detector conv stack as GEMMs (reference
// detection/modules/conv1d.py:38-42).  A 1x1 Conv1d over pooled frames is
//     D[row][n] = sum_k A[row][k] * B[n][k]
// with A = activations, channels-last [rows][K] (rows = clip-major pooled frames,
// each clip padded to a multiple of 128 rows), B = weights [N][K] (PyTorch's own
// (C_out, C_in) layout forward; the transposed copy for the input-gradient GEMM).
//
// k_gemm_tc: tcgen05 tensor cores, kind::tf32, operands staged by TMA into
//   128B-swizzled shared tiles, fp32 accumulator in TMEM, one 128 x BN tile per
//   CTA, warp-specialised (TMA producer / MMA issuer / 4 epilogue warps).
//   The epilogue fuses the InstanceNorm statistics (forward) or the
//   LeakyReLU'/InstanceNorm-adjoint statistics (backward) so the big activations
//   are touched once.
// k_gemm_exact: fp32 CUDA-core GEMM with the same epilogues (precision mode
//   "fp32": used to separate tensor-core rounding from logic errors in the
//   parity tests, and for detection at margins below TF32 resolution).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace aw {

enum { EPI_PLAIN = 0, EPI_FWD = 1, EPI_BWD = 2 };

struct EpiArgs {
  float* out;            // [rows][ldo]
  int ldo;
  int n_valid;           // columns < n_valid are stored
  float* part;           // [row_tiles][ldp][2] per-tile column partial sums (FWD/BWD)
  int ldp;
  const float* act;      // BWD: P_{l-1} [rows][ldo] (post-LeakyReLU activations)
};

// ------------------------------ PTX wrappers --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug traps (-> CUDA error at the next sync) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {
      printf("aware_b200: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128B-swizzled shared tile descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) = 1 | SBO>>4 [32,46) = 1024>>4 | version=1 [46,48)
// | layout_type = SWIZZLE_128B (2) [61,64)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Column sums of a 32(lanes = rows) x 32(registers = columns) block in 31 shuffles:
// after the call lane l holds the sum over the warp's 32 rows of column l.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float keep = up ? v[i + h] : v[i];
      const float send = up ? v[i] : v[i + h];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}

#define AW_GEMM_BK 32                     // 32 fp32 = one 128-byte swizzle row
#define AW_GEMM_STAGES 4

template <int BN>
constexpr int gemm_tc_smem() {
  return AW_GEMM_STAGES * (128 * 128 + BN * 128) + 1024 /*align*/ + 256 /*barriers*/ +
         2 * 4 * BN * 4 /*column partials*/;
}

// grid = (row_tiles, N / BN), block = 192 threads:
//   warp 0: TMA producer, warp 1: MMA issuer (+TMEM alloc), warps 2..5: epilogue.
template <int BN, int EPI>
__global__ void __launch_bounds__(192, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
          int K, EpiArgs ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128, STAGE = A_BYTES + B_BYTES;
  uint8_t* tiles = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + AW_GEMM_STAGES * STAGE);
  uint64_t* empty = full + AW_GEMM_STAGES;
  uint64_t* acc_full = empty + AW_GEMM_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* s_part = reinterpret_cast<float*>(smem + AW_GEMM_STAGES * STAGE + 256);  // [2][4][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_tile = blockIdx.x, n0 = blockIdx.y * BN;
  const int nkb = K / AW_GEMM_BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < AW_GEMM_STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "n"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % AW_GEMM_STAGES;
        const uint32_t ph = (kb / AW_GEMM_STAGES) & 1;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, STAGE);
        tma_load_2d(tiles + s * STAGE, &map_a, full + s, kb * AW_GEMM_BK, row_tile * 128);
        tma_load_2d(tiles + s * STAGE + A_BYTES, &map_b, full + s, kb * AW_GEMM_BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2,
      // K-major A/B, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % AW_GEMM_STAGES;
        const uint32_t ph = (kb / AW_GEMM_STAGES) & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint64_t ad = make_sw128_desc(smem_u32(tiles + s * STAGE));
        const uint64_t bd = make_sw128_desc(smem_u32(tiles + s * STAGE + A_BYTES));
#pragma unroll
        for (int k = 0; k < AW_GEMM_BK / 8; ++k) {
          // advance 8 tf32 = 32 bytes inside the swizzle row: +2 in the (>>4) address field
          tc_mma_tf32(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc,
                      (kb | k) != 0);
        }
        tc_commit(empty + s);
      }
      tc_commit(acc_full);
    }
  } else {
    // ------------------------------ epilogue --------------------------------
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int row = row_tile * 128 + q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* orow = ep.out + (long long)row * ep.ldo;
    const float* arow = EPI == EPI_BWD ? ep.act + (long long)row * ep.ldo : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      const int col = n0 + c0;
      if (EPI == EPI_BWD) {
        // d(IN out) = dP * LeakyReLU'(P);  IN out recovered from P
        float hh[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 p = *reinterpret_cast<const float4*>(arow + col + i);
          const float pp[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool pos = pp[j] > 0.f;
            v[i + j] = pos ? v[i + j] : AW_LEAKY * v[i + j];
            hh[i + j] = (pos ? pp[j] : pp[j] * (1.0f / AW_LEAKY)) * v[i + j];
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          if (col + i < ep.n_valid)
            *reinterpret_cast<float4*>(orow + col + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        const float s1 = warp_colsum32(v, lane);
        const float s2 = warp_colsum32(hh, lane);
        s_part[(0 * 4 + q) * BN + c0 + lane] = s1;
        s_part[(1 * 4 + q) * BN + c0 + lane] = s2;
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          if (col + i < ep.n_valid)
            *reinterpret_cast<float4*>(orow + col + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        if (EPI == EPI_FWD) {
          float sq[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) sq[i] = v[i] * v[i];
          const float s1 = warp_colsum32(v, lane);
          const float s2 = warp_colsum32(sq, lane);
          s_part[(0 * 4 + q) * BN + c0 + lane] = s1;
          s_part[(1 * 4 + q) * BN + c0 + lane] = s2;
        }
      }
    }
    if (EPI != EPI_PLAIN) {
      asm volatile("bar.sync 1, 128;" ::: "memory");   // epilogue warps only
      const int t = threadIdx.x - 64;                  // 0..127
      for (int c = t; c < BN; c += 128) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          s1 += s_part[(0 * 4 + w) * BN + c];
          s2 += s_part[(1 * 4 + w) * BN + c];
        }
        float* p = ep.part + ((long long)row_tile * ep.ldp + n0 + c) * 2;
        p[0] = s1;
        p[1] = s2;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------
// exact fp32 GEMM (CUDA cores) + standalone epilogue with the same semantics
// ---------------------------------------------------------------------------
// D[rows][ldo] = A[rows][K] * B[N][K]^T ; 64x64 tile, 256 threads, 4x4 per thread.
// Products and sums are carried in float64 so the result is the correctly rounded fp32 of
// the exact dot product (better than any fp32 summation order): this path is the
// validation yardstick, not the fast path.
__global__ void __launch_bounds__(256) k_gemm_exact(const float* __restrict__ A,
                                                    const float* __restrict__ B, int K, int N,
                                                    float* __restrict__ D, int ldo) {
  __shared__ float sa[16][64 + 1], sb[16][64 + 1];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, k = i & 15;
      sa[k][r] = A[(long long)(r0 + r) * K + k0 + k];
      sb[k][r] = (c0 + r < N) ? B[(long long)(c0 + r) * K + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sa[k][ty * 4 + i]; b[i] = sb[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c < N) D[(long long)(r0 + ty * 4 + i) * ldo + c] = (float)acc[i][j];
    }
}

// Per 128-row tile column statistics (and the BWD transform), for the exact path.
// grid = (row_tiles, ceil(ncols/128)), block = 128 (one column per thread).
template <int EPI>
__global__ void __launch_bounds__(128) k_epilogue_exact(EpiArgs ep, int ncols) {
  const int row_tile = blockIdx.x;
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= ncols) return;
  float s1 = 0.f, s2 = 0.f;
  for (int r = 0; r < 128; ++r) {
    const long long o = (long long)(row_tile * 128 + r) * ep.ldo + c;
    float v = ep.out[o];
    if (EPI == EPI_BWD) {
      const float p = ep.act[o];
      const bool pos = p > 0.f;
      v = pos ? v : AW_LEAKY * v;
      ep.out[o] = v;
      s1 += v;
      s2 += (pos ? p : p * (1.0f / AW_LEAKY)) * v;
    } else {
      s1 += v;
      s2 += v * v;
    }
  }
  float* p = ep.part + ((long long)row_tile * ep.ldp + c) * 2;
  p[0] = s1;
  p[1] = s2;
}

}  // namespace aw
